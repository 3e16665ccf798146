// Test-only stand-in header: everything lives in opencv2/core/core.hpp (see there).
#pragma once
#include <opencv2/core/core.hpp>
