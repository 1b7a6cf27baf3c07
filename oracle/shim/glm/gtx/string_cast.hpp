// TEST INFRASTRUCTURE ONLY (oracle build shim): forwards to the minimal glm.hpp stand-in.
#pragma once
#include "../glm.hpp"
