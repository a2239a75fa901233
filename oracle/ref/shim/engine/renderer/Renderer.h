// Build shim: the ray marcher only reads `Vulkan.SwapchainExtent.{width,height}`
// (reference RayMarcher.cpp:86-89).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstdint>

struct RefExtentStub
{
	struct { uint32_t width, height; } SwapchainExtent{ 0, 0 };
	static RefExtentStub& GetInstance();   // defined in oracle/ref/harness.cpp
};

#define Vulkan RefExtentStub::GetInstance()
