// stand-in for boost::hash_combine (point_cloud_utils.cpp:18)
#pragma once
#include <cstddef>
namespace boost { inline void hash_combine(std::size_t& seed, std::size_t v) { seed ^= v + 0x9e3779b9 + (seed << 6) + (seed >> 2); } }
