// Internal helpers shared by api.cpp and plan.cpp.
#pragma once
#include "../../include/pvsr.h"
#include <vector>

#include "conv.h"
#include "wgrad.h"

namespace pvsr {
int set_error(int code, const char* fmt, ...);
int check_cuda(int cuda_error, const char* what);
int device_num_sms();
int fill_conv_params(const pvsr_conv_desc* d, ConvParams* p);

// One X source of a weight-gradient problem (its taps / channel blocks become units) and one dY chunk.
struct WgSource { SrcView view; };
struct WgChunk { SrcView view; int col0; };
// Splits (units x chunks) into CTA-pair jobs; the optional bias ("ones") unit rides in the last unit group.
// contig_cols > 0: chunks[0] is the first view of one tensor with that many consecutive gradient channels.
void build_wgrad_jobs(const std::vector<WgSource>& srcs, int kb_per_src, int taps, const std::vector<WgChunk>& chunks,
                      int n_total, bool with_bias, long long dw_off, long long db_off, std::vector<WgJob>* out,
                      int contig_cols = 0);
}  // namespace pvsr
