// Build shim for compiling the reference's hot-path translation units on Linux/g++.
// Replaces /root/reference/src/engine/hzpch.h (the precompiled header), which pulls in
// Vulkan, GLFW, spdlog, yaml-cpp and ImGui -- none of which the CPU ray-march path uses.
// TEST INFRASTRUCTURE ONLY (see oracle/README.md).  The three GLM_FORCE_* defines are part
// of the numeric contract (reference hzpch.h:22-24) and are kept verbatim in meaning.
#pragma once

#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <functional>
#include <iostream>
#include <list>
#include <memory>
#include <optional>
#include <queue>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#define GLM_FORCE_RADIANS
#define GLM_FORCE_DEPTH_ZERO_TO_ONE
#define GLM_FORCE_LEFT_HANDED

#include <glm/glm.hpp>
#include <glm/vec4.hpp>
#include <glm/mat4x4.hpp>
#include <glm/gtc/quaternion.hpp>
#include <glm/ext/matrix_transform.hpp>
#include <glm/ext/matrix_clip_space.hpp>

#include <Eigen/Dense>

// logging / assertion macros of the engine: compiled out
#define SPDLOG_TRACE(...) ((void)0)
#define SPDLOG_DEBUG(...) ((void)0)
#define SPDLOG_INFO(...) ((void)0)
#define SPDLOG_WARN(...) ((void)0)
#define SPDLOG_ERROR(...) ((void)0)
#define SPDLOG_CRITICAL(...) ((void)0)
#define HZ_ASSERT(cond, ...) ((void)0)

// GLFW button ids used by the orbit controller's event handler
#define GLFW_MOUSE_BUTTON_LEFT 0
#define GLFW_MOUSE_BUTTON_RIGHT 1
