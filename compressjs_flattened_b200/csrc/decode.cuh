// decode.cuh -- decompress kernels (K-U1..U4); filled in below.
#pragma once
#include "common.cuh"
