// TEST INFRASTRUCTURE ONLY: forwards to the API stubs (see ros_stub_core.hpp).
#pragma once
#include "../../ros_stub_core.hpp"
