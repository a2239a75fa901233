// reference_caller_check.cpp -- compile-only proof that host/RayMarcher.h is a drop-in for the reference's
// RayMarcher.h: this translation unit includes the REFERENCE's own engine/app headers (Camera3D,
// CameraController3D, Dataset) and repeats the call block of AdvancedRenderer::Render
// (src/app/AdvancedRenderer/AdvancedRenderer.cpp:257-298, 129) against our class.  Built with -fsyntax-only by
// tests/test_cabi.py when /root/reference is present; never linked, never shipped.
#include "engine/hzpch.h"                       // oracle/ref/shim stand-in for the precompiled header
#include <engine/camera/Camera3D.h>
#include <engine/camera/CameraController3D.h>
#include <engine/renderer/Renderer.h>           // oracle/ref/shim: the `Vulkan.SwapchainExtent` the marcher reads
#include "app/Dataset.h"

#include "RayMarcher.h"                         // ours

static VisualizationSettings g_VisualizationSettings = {
	.Frame = 0, .MaxSteps = 128, .StepSize = 0.009f, .IsoDensity = 1.0f, .EnableAnisotropy = false,
	.k_n = 0.5f, .k_r = 2.0f, .k_s = 2000.0f, .N_eps = 1,
};

struct CallerLikeAdvancedRenderer
{
	Camera3D Camera;
	CameraController3D CameraController{ Camera };
	::Dataset* Dataset = nullptr;
	RayMarcher m_RayMarcher;
	glm::vec4* m_Positions = nullptr;
	glm::vec4* m_Normals = nullptr;
	float* m_Depth = nullptr;
	bool RayMarchFinished = true;

	void Render()
	{
		if (RayMarchFinished)
		{
			RayMarchFinished = false;
			m_RayMarcher.Prepare(g_VisualizationSettings, CameraController, Dataset, m_Positions, m_Normals, m_Depth);
			m_RayMarcher.Start();
		}
		if (!RayMarchFinished && m_RayMarcher.IsDone())
		{
			RayMarchFinished = true;
			g_VisualizationSettings.Frame++;
		}
	}

	void Exit() { m_RayMarcher.Exit(); }
};

void instantiate(CallerLikeAdvancedRenderer& r) { r.Render(); r.Exit(); }

// the call block as host/engine.patch leaves it (no host round trip): the members engine.patch adds to BilateralBuffer are
// stood in for by this stub -- no Vulkan headers exist in this image, so the Vulkan side of the patch is not compiled
struct ExportableBufferStub
{
	bool Exportable = false;
	size_t AllocationSize = 0;
	int ExportFd() { return -1; }
	void CopyToGPU() {}
};

struct PatchedAdvancedRenderer
{
	Camera3D Camera;
	CameraController3D CameraController{ Camera };
	::Dataset* Dataset = nullptr;
	RayMarcher m_RayMarcher;
	ExportableBufferStub PositionsBuffer, NormalsBuffer;
	bool RayMarchFinished = true;

	void Init()
	{
		PositionsBuffer.Exportable = true;
		NormalsBuffer.Exportable = true;
		m_RayMarcher.SetOutput(RayMarcher::Output::DeviceOnly);
		m_RayMarcher.SetUseGpuDepthPrePass(true);
		m_RayMarcher.Prepare(g_VisualizationSettings, CameraController, Dataset, nullptr, nullptr, nullptr);
		if (fr_import_vk_images_fd(m_RayMarcher.Context(), PositionsBuffer.ExportFd(), NormalsBuffer.ExportFd(),
								   PositionsBuffer.AllocationSize) != FR_OK)
			SPDLOG_ERROR("fluidmarch: {}", fr_last_error());
	}

	void Render()
	{
		if (RayMarchFinished)
		{
			RayMarchFinished = false;
			m_RayMarcher.Prepare(g_VisualizationSettings, CameraController, Dataset, nullptr, nullptr, nullptr);
			m_RayMarcher.Start();
		}
		if (!RayMarchFinished && m_RayMarcher.IsDone()) RayMarchFinished = true;
		PositionsBuffer.CopyToGPU();
		NormalsBuffer.CopyToGPU();
	}
};

void instantiate_patched(PatchedAdvancedRenderer& r) { r.Init(); r.Render(); }
