// RayMarcher.h -- header-compatible B200 replacement of the reference's CPU ray marcher class
// (src/app/AdvancedRenderer/RayMarcher.h:12-44, RayMarcher.cpp:64-112) over the C ABI of libfluidmarch.so
// (include/fluidmarch.h).  Put this directory before src/ on the include path (or copy the file over
// src/app/AdvancedRenderer/RayMarcher.h) and drop RayMarcher.cpp / ThreadPool.cpp from the build:
// AdvancedRenderer.cpp compiles unchanged -- same type names, same member functions, same call protocol
//
//     m_RayMarcher.Prepare(settings, CameraController, Dataset, positions, normals, depth);
//     m_RayMarcher.Start();                 // returns immediately            (AdvancedRenderer.cpp:265-272)
//     if (m_RayMarcher.IsDone()) { ... }    // polled once per UI frame       (AdvancedRenderer.cpp:275)
//     m_RayMarcher.Exit();                  // at shutdown                    (AdvancedRenderer.cpp:129)
//
// Everything the reference's engine types contribute is read through their public members, exactly the ones
// RayMarcher.cpp reads (camera.Position, camera.Camera.GetInvProjectionView(), dataset->ParticleRadius,
// dataset->Frames[Frame].m_Particles, Vulkan.SwapchainExtent), so the header needs no engine include of its own:
// Prepare is a template over the caller's types.  There is no CPU fallback: if the library or a B200 is missing,
// Prepare reports the error through FLUIDMARCH_ERROR (SPDLOG_ERROR inside the engine) and IsDone() stays true
// with zeroed outputs -- the same "errors are logged, nothing is thrown" contract the reference has
// (SURVEY.md 8b).
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "fluidmarch.h"

#ifndef FLUIDMARCH_ERROR
#ifdef SPDLOG_ERROR
#define FLUIDMARCH_ERROR(msg) SPDLOG_ERROR("fluidmarch: {}", msg)
#else
#define FLUIDMARCH_ERROR(msg) std::fprintf(stderr, "fluidmarch: %s\n", msg)
#endif
#endif

class Dataset;

// src/app/AdvancedRenderer/RayMarcher.h:12-26 (defaults live in AdvancedRenderer.cpp:18-28)
struct VisualizationSettings
{
	int Frame;

	int MaxSteps;
	float StepSize;
	float IsoDensity;

	bool EnableAnisotropy;

	float k_n;
	float k_r;
	float k_s;
	int N_eps;
};

class RayMarcher
{
public:
	// how the finished image leaves the GPU
	enum class Output
	{
		HostBuffers,     // reference behaviour: positions / normals are copied into the caller's (host-mapped) buffers
		DeviceOnly       // interop: results stay on the GPU (fr_device_images / fr_import_vk_memory_fd); the
		                 // positions / normals pointers may be null
	};

	RayMarcher() = default;
	RayMarcher(const RayMarcher&) = delete;
	RayMarcher& operator=(const RayMarcher&) = delete;
	~RayMarcher() { Exit(); }

	// RayMarcher::Exit (RayMarcher.cpp:71-74)
	void Exit()
	{
		if (m_Ctx) fr_destroy(m_Ctx);
		m_Ctx = nullptr;
		m_Resident.clear();
		m_Running = false;
	}

	// knobs that do not exist in the reference; defaults reproduce it
	void SetDevice(int device) { m_Device = device; }
	void SetExtent(uint32_t width, uint32_t height) { m_Width = width; m_Height = height; }   // instead of Vulkan.SwapchainExtent
	void SetOutput(Output o) { m_Output = o; }
	void SetBisectionSteps(int n) { m_BisectionSteps = n; }           // 0 = hit is the first sample >= iso, as in the reference
	void SetUseGpuDepthPrePass(bool on) { m_GpuDepth = on; }           // true: `depth` is ignored, the CUDA pre-pass makes it
	void SetFastNormals(bool on) { m_FastNormals = on; }               // see fr_settings::fast_normals
	void SetSkipLastPixel(bool on) { m_SkipLastPixel = on; }           // ThreadPool.cpp:50 leaves pixel W*H-1 untouched
	fr_context* Context() { return m_Ctx; }

	// interop call (host/engine.patch): no host buffers at all -- Output::DeviceOnly with the GPU depth pre-pass
	struct Float4 { float x, y, z, w; };
	template <class Controller, class DatasetT>
	void Prepare(const VisualizationSettings& settings, const Controller& camera, DatasetT* dataset, std::nullptr_t, std::nullptr_t,
				 std::nullptr_t)
	{
		Prepare(settings, camera, dataset, static_cast<Float4*>(nullptr), static_cast<Float4*>(nullptr), static_cast<float*>(nullptr));
	}

	// RayMarcher::Prepare (RayMarcher.cpp:76-100).  Controller = CameraController3D, DatasetT = Dataset,
	// Vec4 = glm::vec4 (any 16-byte POD of four floats).
	template <class Controller, class DatasetT, class Vec4>
	void Prepare(const VisualizationSettings& settings, const Controller& camera, DatasetT* dataset, Vec4* positions,
				 Vec4* normals, float* depth)
	{
		static_assert(sizeof(Vec4) == 16, "positions / normals must be 4 x float images");
		m_Ready = false;
		m_Positions = reinterpret_cast<float*>(positions);
		m_Normals = reinterpret_cast<float*>(normals);
		uint32_t w = m_Width, h = m_Height;
#ifdef Vulkan
		if (!w || !h) { w = Vulkan.SwapchainExtent.width; h = Vulkan.SwapchainExtent.height; }   // RayMarcher.cpp:86-89
#endif
		if (!w || !h) return Fail("no extent: call SetExtent() (no `Vulkan` macro in this translation unit)");
		if (!dataset) return Fail("null dataset");
		if (settings.Frame < 0 || (size_t)settings.Frame >= dataset->Frames.size()) return Fail("settings.Frame is not a frame of the dataset");
		if (!EnsureContext(w, h)) return;

		// frame -> GPU (once per frame index): Frame::m_Particles is a packed array of glm::vec3
		auto& frame = dataset->Frames[(size_t)settings.Frame];
		static_assert(sizeof(frame.m_Particles[0]) == 12, "Particle must be a packed float3");
		const float* xyz = reinterpret_cast<const float*>(frame.m_Particles.data());
		size_t const n = frame.m_Particles.size();
		auto it = m_Resident.find(settings.Frame);
		if (it == m_Resident.end() || it->second.first != xyz || it->second.second != n)
		{
			float const mult = dataset->ParticleRadius != 0.0f ? dataset->ParticleRadiusExt / dataset->ParticleRadius : 2.0f;
			if (fr_upload_frame(m_Ctx, settings.Frame, xyz, n, dataset->ParticleRadius, mult) != FR_OK) return Fail(fr_last_error());
			m_Resident[settings.Frame] = { xyz, n };
		}

		fr_settings s;
		std::memset(&s, 0, sizeof s);
		s.frame = settings.Frame;
		s.max_steps = settings.MaxSteps;
		s.step_size = settings.StepSize;
		s.iso_density = settings.IsoDensity;
		s.enable_anisotropy = settings.EnableAnisotropy ? 1 : 0;
		s.k_n = settings.k_n; s.k_r = settings.k_r; s.k_s = settings.k_s; s.n_eps = settings.N_eps;
		s.bisection_steps = m_BisectionSteps;
		s.skip_last_pixel = m_SkipLastPixel ? 1 : 0;
		s.fast_normals = m_FastNormals ? 1 : 0;
		if (fr_set_settings(m_Ctx, &s) != FR_OK) return Fail(fr_last_error());

		// camera: the members RayMarcher.cpp:95-96 reads, plus View / Projection / System[2] for the depth
		// pre-pass and the shading (DepthRenderPass.cpp:100-108, CompositionRenderPass.cpp:313-321)
		fr_camera c;
		std::memcpy(c.view, &camera.Camera.GetView(), 64);
		std::memcpy(c.projection, &camera.Camera.GetProjection(), 64);
		std::memcpy(c.inv_projection_view, &camera.Camera.GetInvProjectionView(), 64);
		std::memcpy(c.position, &camera.Position, 12);
		std::memcpy(c.direction, reinterpret_cast<const float*>(&camera.System) + 6, 12);   // glm::mat3 column 2
		if (fr_set_camera(m_Ctx, &c) != FR_OK) return Fail(fr_last_error());

		if (!m_GpuDepth)
		{
			if (!depth) return Fail("null depth image");
			if (fr_set_depth(m_Ctx, depth) != FR_OK) return Fail(fr_last_error());
		}
		m_Ready = true;
	}

	// RayMarcher::Start (RayMarcher.cpp:102-112): returns immediately, the frame renders on the context's stream
	void Start()
	{
		m_Running = false;
		if (!m_Ready) return ZeroOutputs();
		int const passes = (m_GpuDepth ? FR_PASS_DEPTH : 0) | FR_PASS_MARCH | FR_PASS_SHADE;
		if (fr_render_async(m_Ctx, passes) != FR_OK) { Fail(fr_last_error()); return ZeroOutputs(); }
		m_Running = true;
	}

	// RayMarcher::IsDone (RayMarcher.h:44).  Outputs are valid once this returned true.
	bool IsDone()
	{
		if (!m_Running) return true;
		int const rc = fr_is_done(m_Ctx);
		if (rc == 0) return false;
		m_Running = false;
		if (rc < 0) { Fail(fr_last_error()); ZeroOutputs(); return true; }
		if (m_Output == Output::HostBuffers && fr_download(m_Ctx, nullptr, m_Positions, m_Normals, nullptr) != FR_OK)
		{
			Fail(fr_last_error());
			ZeroOutputs();
		}
		return true;
	}

	const char* LastError() const { return m_Error.empty() ? nullptr : m_Error.data(); }

private:
	bool EnsureContext(uint32_t w, uint32_t h)
	{
		if (m_Ctx && (w != m_CtxW || h != m_CtxH))
		{
			if (fr_resize(m_Ctx, (int)w, (int)h) != FR_OK) { Fail(fr_last_error()); return false; }
		}
		else if (!m_Ctx)
		{
			if (fr_create(m_Device, (int)w, (int)h, &m_Ctx) != FR_OK) { m_Ctx = nullptr; Fail(fr_last_error()); return false; }
			fr_set_stage_timing(m_Ctx, 0);      // nobody reads fr_get_timings here: no events between the kernels of a frame
		}
		m_CtxW = w; m_CtxH = h;
		return true;
	}

	void Fail(const char* msg)
	{
		m_Error.assign(msg, msg + std::strlen(msg) + 1);
		FLUIDMARCH_ERROR(msg);
	}

	void ZeroOutputs()
	{
		if (m_Output != Output::HostBuffers || !m_CtxW) return;
		size_t const bytes = (size_t)m_CtxW * m_CtxH * 16;
		if (m_Positions) std::memset(m_Positions, 0, bytes);
		if (m_Normals) std::memset(m_Normals, 0, bytes);
	}

	fr_context* m_Ctx = nullptr;
	int m_Device = 0;
	uint32_t m_Width = 0, m_Height = 0, m_CtxW = 0, m_CtxH = 0;
	Output m_Output = Output::HostBuffers;
	int m_BisectionSteps = 0;
	bool m_GpuDepth = false, m_SkipLastPixel = false, m_FastNormals = false;
	bool m_Ready = false, m_Running = false;
	float* m_Positions = nullptr;
	float* m_Normals = nullptr;
	std::unordered_map<int, std::pair<const float*, size_t>> m_Resident;   // frame index -> uploaded array
	std::vector<char> m_Error;
};
