// Same include path as cilantro's core/grid_downsampler.hpp; the B200-native drop-in lives in b200_shims.hpp.
#pragma once
#include "../b200_shims.hpp"
