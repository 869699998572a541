// TEST INFRASTRUCTURE (oracle build only): forwards <nameof/nameof.hpp> to the
// vendored header, whose real location is external/nameof/nameof/include.
#pragma once
#include "nameof/include/nameof.hpp"
