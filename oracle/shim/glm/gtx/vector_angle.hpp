// TEST INFRASTRUCTURE ONLY (oracle build shim): glm/gtx/vector_angle.hpp
// (included by advect_floating_items.cpp:6; nothing from it is used there).
#pragma once
#include "../glm.hpp"
