// Error state, version and small host-side helpers of the C ABI (include/ngp_b200.h).
#include "ngp_common.cuh"
#include <vector>
#include <utility>
#include <mutex>

namespace {
std::mutex g_err_mu;
thread_local std::string g_err;
}  // namespace

void ngp_set_error(const std::string& msg) { g_err = msg; }

int ngp_num_sms() {
    static int sms[64] = {0};                                     // per device: a process may drive several GPUs
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (sms[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
        sms[dev] = n;
    }
    return sms[dev];
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device) instead of on every launch
bool ngp_first_use(const void* kernel) {
    static std::vector<std::pair<const void*, int>> seen;
    int dev = 0;
    cudaGetDevice(&dev);
    for (const auto& e : seen)
        if (e.first == kernel && e.second == dev) return false;
    seen.emplace_back(kernel, dev);
    return true;
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
