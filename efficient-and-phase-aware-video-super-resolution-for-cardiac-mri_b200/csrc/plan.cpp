// placeholder until the plan runtime lands (next commit)
#include "internal.h"
using namespace pvsr;
struct pvsr_plan { pvsr_net_config cfg; };
extern "C" {
int pvsr_plan_create(const pvsr_net_config*, pvsr_plan**) { return set_error(-1, "not implemented"); }
void pvsr_plan_destroy(pvsr_plan*) {}
int64_t pvsr_plan_workspace_bytes(const pvsr_plan*) { return 0; }
int64_t pvsr_plan_packed_bytes(const pvsr_plan*) { return 0; }
int64_t pvsr_plan_output_elems(const pvsr_plan*) { return 0; }
int pvsr_plan_num_lists(const pvsr_plan*) { return 0; }
int64_t pvsr_plan_num_launches(const pvsr_plan*) { return 0; }
double pvsr_plan_flops(const pvsr_plan*) { return 0; }
int pvsr_plan_pack(pvsr_plan*, const pvsr_net_params*, void*, void*) { return set_error(-1, "not implemented"); }
int pvsr_plan_forward(pvsr_plan*, const pvsr_net_params*, const void*, const float*, const float*, float*, void*, int,
                      void*) { return set_error(-1, "not implemented"); }
}
