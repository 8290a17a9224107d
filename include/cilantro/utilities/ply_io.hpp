// Same include path as cilantro's utilities/ply_io.hpp; the PLY passthrough lives in b200_ply.hpp (used by PointCloud3f).
#pragma once
#include "../b200_shims.hpp"
