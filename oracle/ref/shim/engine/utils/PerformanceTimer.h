// Build shim: profiling scopes compiled out.  TEST INFRASTRUCTURE ONLY.
#pragma once
#define PROFILE_SCOPE(name) ((void)0)
#define PROFILE_FUNCTION() ((void)0)
