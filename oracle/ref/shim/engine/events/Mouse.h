// Build shim: mouse event payloads read by CameraController3D::HandleEvent.
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include "engine/events/EventManager.h"
#include <glm/glm.hpp>

namespace Events
{
	struct MouseDown : Event { int Button = 0; };
	struct MouseUp : Event { int Button = 0; };
	struct MouseWheel : Event { float Wheel = 0.0f; };
	struct MouseMove : Event { glm::vec2 Delta{ 0.0f }; };
}
