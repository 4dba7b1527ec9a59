// Device-resident scan (internal header): n records of `stride` bytes, laid out like mimosa::lidar::Point
// (mimosa/include/mimosa/lidar/point.hpp:18-39; xyz are the first three floats).  Lets the steps either side
// of the ICP factor — deskew, body-frame transform, voxel downsample, map insertion — run without host
// round-trips.
#pragma once
#include "mb_internal.cuh"

struct mb_scan {
  mb_ctx* ctx = nullptr;
  unsigned char* data = nullptr;  // pooled device block
  size_t n = 0, stride = 0, bytes = 0;
};
