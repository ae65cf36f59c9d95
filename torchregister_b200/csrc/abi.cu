// abi.cu — error plumbing and version entry points of the C ABI (include/trb.h).
#include "common.cuh"
#include <stdarg.h>
#include <stdio.h>

namespace trb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return TRB_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}

}  // namespace trb

extern "C" int trb_abi_version(void) { return TRB_ABI_VERSION; }
extern "C" const char *trb_last_error(void) { return trb::g_err; }
