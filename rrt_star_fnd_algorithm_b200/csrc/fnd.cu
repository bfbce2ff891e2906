// fnd.cu — RRT*-FND tree surgery (placeholder while the kernels are being written; replaced below).
#include <cuda_runtime.h>
#include "../../include/rrtfnd.h"
extern "C" int rrtfnd_fnd_select_branch(const rrtfnd_params_t*, const rrtfnd_batch_t*, const int32_t*, void*) { return RRTFND_E_BADARG; }
extern "C" int rrtfnd_fnd_valid_path(const rrtfnd_params_t*, const rrtfnd_batch_t*, int32_t*, void*) { return RRTFND_E_BADARG; }
extern "C" int rrtfnd_fnd_reconnect(const rrtfnd_params_t*, const rrtfnd_batch_t*, const int32_t*, double, int32_t*, void*) { return RRTFND_E_BADARG; }
