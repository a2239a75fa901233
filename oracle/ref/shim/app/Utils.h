// Build shim for src/app/Utils.h: only RandomFloat is needed (Dataset::makeCube).
// Deterministic here on purpose.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <random>

inline float RandomFloat(float lo, float hi)
{
	static std::mt19937 gen(20240229u);
	std::uniform_real_distribution<float> d(lo, hi);
	return d(gen);
}
