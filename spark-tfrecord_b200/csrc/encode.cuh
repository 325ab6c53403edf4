// encode.cuh -- K5: columnar rows -> protobuf wire bytes -> CRC framing (mirror of decode).
#pragma once
#include "common.cuh"
