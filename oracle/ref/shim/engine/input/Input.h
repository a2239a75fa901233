// Build shim: nothing from the input layer is used on the hot path.  TEST INFRASTRUCTURE ONLY.
#pragma once
