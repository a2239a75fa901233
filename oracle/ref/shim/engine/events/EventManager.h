// Build shim: minimal event dispatch so CameraController3D.cpp compiles unmodified.
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include <typeinfo>

class Event
{
public:
	virtual ~Event() = default;
};

class EventDispatcher
{
public:
	explicit EventDispatcher(Event& e) : m_Event(e) {}

	template <typename T, typename F>
	bool Dispatch(const F& f)
	{
		if (auto* p = dynamic_cast<T*>(&m_Event)) { f(*p); return true; }
		return false;
	}
private:
	Event& m_Event;
};
