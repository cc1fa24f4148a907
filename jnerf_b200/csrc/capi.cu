// Error state, version and small host-side helpers of the C ABI (include/ngp_b200.h).
#include "ngp_common.cuh"
#include <mutex>

namespace {
std::mutex g_err_mu;
thread_local std::string g_err;
}  // namespace

void ngp_set_error(const std::string& msg) { g_err = msg; }

int ngp_num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
        sms = p.multiProcessorCount;
    }
    return sms;
}

extern "C" {

const char* ngp_last_error(void) { return g_err.c_str(); }
int ngp_version(void) { return 100; }
int ngp_sm_count(void) { return ngp_num_sms(); }

// ops/op_include/pcg32/pcg32.h:53-60 (seed) and :145-166 (advance); pure integer, host side
void ngp_pcg32_seed(uint64_t initstate, uint64_t initseq, uint64_t* state_inc) {
    Pcg32 r;
    r.state = 0;
    r.inc = (initseq << 1u) | 1u;
    r.next_uint();
    r.state += initstate;
    r.next_uint();
    state_inc[0] = r.state;
    state_inc[1] = r.inc;
}
void ngp_pcg32_advance(uint64_t* state_inc, int64_t delta) {
    Pcg32 r{state_inc[0], state_inc[1]};
    r.advance(delta);
    state_inc[0] = r.state;
}

}  // extern "C"
