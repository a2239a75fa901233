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
