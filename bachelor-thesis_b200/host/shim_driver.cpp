// shim_driver.cpp -- drives host/RayMarcher.h exactly the way AdvancedRenderer::Render drives the reference's
// RayMarcher (src/app/AdvancedRenderer/AdvancedRenderer.cpp:257-298), with stand-ins for the engine types that
// have the same member names the marcher reads.  Used by tests/test_gpu_parity.py::test_cpp_shim_*:
//
//     shim_driver <in.bin> <out.bin> [gpu_depth]
//
// in.bin : int32 W, H, n, frame_count | VisualizationSettings | view[16] proj[16] ipv[16] position[3] system[9]
//          | float h, mult | n x float3 particles | W*H float depth
// out.bin: W*H*4 float positions | W*H*4 float normals | int32 polls
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "RayMarcher.h"

struct Mat4 { float m[16]; };
struct Mat3 { float m[9]; };
struct Vec3 { float x, y, z; };
struct Vec4 { float x, y, z, w; };

// src/engine/camera/Camera3D.h:13-32
struct Camera3D
{
	const Mat4& GetProjection() const { return Projection; }
	const Mat4& GetView() const { return View; }
	const Mat4& GetInvProjectionView() const { return InvProjectionView; }
	Mat4 Projection, View, InvProjectionView;
};

// src/engine/camera/CameraController3D.h:7-27
struct CameraController3D
{
	explicit CameraController3D(Camera3D& camera) : Camera(camera) {}
	Camera3D& Camera;
	Vec3 Position;
	Mat3 System;
};

// src/app/Dataset.h:37-105
struct Frame { std::vector<Vec3> m_Particles; };
struct DatasetStandIn
{
	float ParticleRadius, ParticleRadiusExt;
	std::vector<Frame> Frames;
};

static void rd(FILE* f, void* p, size_t bytes)
{
	if (fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "shim_driver: short read\n"); exit(2); }
}

int main(int argc, char** argv)
{
	if (argc < 3) { fprintf(stderr, "usage: shim_driver in.bin out.bin [gpu_depth]\n"); return 2; }
	FILE* in = fopen(argv[1], "rb");
	if (!in) { perror(argv[1]); return 2; }
	int32_t hdr[4];
	rd(in, hdr, sizeof hdr);
	int const W = hdr[0], H = hdr[1], n = hdr[2];
	VisualizationSettings settings;
	rd(in, &settings, sizeof settings);
	Camera3D cam;
	rd(in, &cam.View, 64); rd(in, &cam.Projection, 64); rd(in, &cam.InvProjectionView, 64);
	CameraController3D controller(cam);
	rd(in, &controller.Position, 12); rd(in, &controller.System, 36);
	float hm[2];
	rd(in, hm, 8);
	DatasetStandIn dataset;
	dataset.ParticleRadius = hm[0];
	dataset.ParticleRadiusExt = hm[0] * hm[1];
	dataset.Frames.resize((size_t)settings.Frame + 1);
	dataset.Frames[(size_t)settings.Frame].m_Particles.resize((size_t)n);
	rd(in, dataset.Frames[(size_t)settings.Frame].m_Particles.data(), (size_t)n * 12);
	std::vector<float> depth((size_t)W * H);
	rd(in, depth.data(), depth.size() * 4);
	fclose(in);

	std::vector<Vec4> positions((size_t)W * H, Vec4{ 9, 9, 9, 9 }), normals((size_t)W * H, Vec4{ 9, 9, 9, 9 });

	RayMarcher marcher;
	marcher.SetExtent((uint32_t)W, (uint32_t)H);          // the engine build reads Vulkan.SwapchainExtent instead
	marcher.SetUseGpuDepthPrePass(argc > 3);
	int polls = 0;
	for (int round = 0; round < 2; round++)                // twice: the second Prepare must reuse the resident frame
	{
		marcher.Prepare(settings, controller, &dataset, positions.data(), normals.data(), depth.data());
		marcher.Start();
		while (!marcher.IsDone())
		{
			polls++;
			std::this_thread::sleep_for(std::chrono::microseconds(50));
		}
	}
	if (marcher.LastError()) { fprintf(stderr, "shim_driver: %s\n", marcher.LastError()); return 1; }
	marcher.Exit();

	FILE* out = fopen(argv[2], "wb");
	if (!out) { perror(argv[2]); return 2; }
	fwrite(positions.data(), 16, positions.size(), out);
	fwrite(normals.data(), 16, normals.size(), out);
	int32_t p32 = polls;
	fwrite(&p32, 4, 1, out);
	fclose(out);
	return 0;
}
