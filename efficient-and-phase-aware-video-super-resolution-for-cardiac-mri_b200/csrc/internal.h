// Internal helpers shared by api.cpp and plan.cpp.
#pragma once
#include "../../include/pvsr.h"
#include "conv.h"

namespace pvsr {
int set_error(int code, const char* fmt, ...);
int check_cuda(int cuda_error, const char* what);
int device_num_sms();
int fill_conv_params(const pvsr_conv_desc* d, ConvParams* p);
}  // namespace pvsr
