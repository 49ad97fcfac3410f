#pragma once
#include <cho_util/core/geometry/point_cloud.hpp>
