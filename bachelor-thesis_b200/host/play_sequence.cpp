// play_sequence.cpp -- the reference's start-up + autoplay + recording loop over the C ABI, in C++:
//   main.cpp:48-53            new Dataset("datasets/<name>/ParticleData_Fluid_", ".bgeo", h, mult, count)
//   AdvancedRenderer.cpp:275-298   wait for the march, Frame++, Vulkan.Screenshot() while g_Recording
//   Renderer.cpp:400-409      screenshots/screenshot_<n>.bmp
// with `lanes` frames in flight on one GPU (fr_seq_*): files are read and decoded, frames rendered and screenshots
// written by the lanes' workers.  Used by tests/test_bgeo.py (GPU) and as the example a maintainer starts from.
//
//     play_sequence <prefix> <suffix> <count|-1> <W> <H> <lanes> <camera.bin> <out_prefix> [h] [mult] [aniso]
//
// camera.bin: fr_camera as 54 little-endian floats (view, projection, inv_projection_view, position, direction) --
// what CameraController3D / Camera3D hold (CameraController3D.h:22-27, Camera3D.h:30-32).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fluidmarch.h"

static int fail(const char* what)
{
	std::fprintf(stderr, "play_sequence: %s: %s\n", what, fr_last_error());
	return 1;
}

int main(int argc, char** argv)
{
	if (argc < 9) { std::fprintf(stderr, "usage: play_sequence prefix suffix count W H lanes camera.bin out_prefix [h] [mult] [aniso]\n"); return 2; }
	std::string const prefix = argv[1], suffix = argv[2], out_prefix = argv[8];
	int const want = std::atoi(argv[3]), W = std::atoi(argv[4]), H = std::atoi(argv[5]), lanes = std::atoi(argv[6]);
	float const h = argc > 9 ? (float)std::atof(argv[9]) : 0.1f;          // particleRadius           (assets/config.yml:19)
	float const mult = argc > 10 ? (float)std::atof(argv[10]) : 2.0f;     // particleRadiusMultiplier (assets/config.yml:20)
	bool const aniso = argc > 11 && std::atoi(argv[11]) != 0;

	fr_camera cam;
	FILE* cf = std::fopen(argv[7], "rb");
	if (!cf || std::fread(&cam, 1, sizeof cam, cf) != sizeof cam) { std::fprintf(stderr, "play_sequence: cannot read %s\n", argv[7]); return 2; }
	std::fclose(cf);

	int const count = fr_dataset_count(prefix.c_str(), suffix.c_str(), want);          // Dataset.cpp:186-203
	if (count <= 0) { std::fprintf(stderr, "play_sequence: no files %s1%s ...\n", prefix.c_str(), suffix.c_str()); return 2; }

	fr_sequence* seq = nullptr;
	if (fr_seq_create(0, W, H, lanes, &seq) != FR_OK) return fail("fr_seq_create");
	fr_settings s;
	std::memset(&s, 0, sizeof s);
	s.max_steps = 128; s.step_size = 0.009f; s.iso_density = 1.0f;                       // AdvancedRenderer.cpp:18-28
	s.enable_anisotropy = aniso ? 1 : 0; s.k_n = 0.5f; s.k_r = 2.0f; s.k_s = 2000.0f; s.n_eps = 1;
	if (fr_seq_set_settings(seq, &s) != FR_OK || fr_seq_set_camera(seq, &cam) != FR_OK) return fail("settings / camera");

	std::vector<std::string> files((size_t)count), shots((size_t)count);
	for (int i = 0; i < count; i++)
	{
		files[(size_t)i] = prefix + std::to_string(i + 1) + suffix;
		shots[(size_t)i] = out_prefix + std::to_string(i) + ".bmp";                         // screenshot_%d.bmp, from 0
		fr_seq_job job;
		std::memset(&job, 0, sizeof job);
		job.bgeo_path = files[(size_t)i].c_str();
		job.bmp_path = shots[(size_t)i].c_str();
		job.h = h; job.h_ext_mult = mult;
		if (fr_seq_submit(seq, &job) < 0) return fail("fr_seq_submit");
	}
	if (fr_seq_drain(seq) != FR_OK) return fail("frame");
	std::printf("played %d frames, %d in flight\n", count, lanes);
	fr_seq_destroy(seq);
	return 0;
}
