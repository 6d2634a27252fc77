#include "llgrad_tc.cuh"

namespace pyglm {

void TcWorkspace::release()
{
    if (buf) cudaFree(buf);
    buf = nullptr; bytes = 0;
    if (tmap) free(tmap);
    tmap = nullptr;
}

bool tc_supported(int64_t, int, int, int) { return false; }

int launch_tc_ll_grad(const TcArgs&, TcWorkspace&, cudaStream_t)
{
    set_error("tensor-core path not built");
    return PYGLM_B200_EUNSUPPORTED;
}

}  // namespace pyglm
