// host_api.cu — host-buffer entry point of the C ABI: rrtfnd_plan_host.
//
// This is what the Python drop-in (RRT / RRT_star / RRT_star_FN with reference signatures) and the
// end-to-end leg of bench.py call: inputs and outputs are HOST buffers, the library owns the device
// memory for the duration of the call, and the host<->device copies are part of the call.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/rrtfnd.h"

namespace {

struct DevBuf {
    void* p = nullptr;
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    ~DevBuf() { if (p) cudaFree(p); }
};

#define CK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return (int)e_; } while (0)

}  // namespace

extern "C" int rrtfnd_plan_host(const rrtfnd_params_t* prm, int32_t P, const double* starts_goals_host,
                                const uint64_t* seeds_host, const double* obst_host, const uint32_t* words_host,
                                rrtfnd_planner_t* planners_host, double* x_host, double* y_host, double* ecost_host,
                                int32_t* parent_host, int32_t* nchild_host, int32_t* id_host, int32_t* best_path_host,
                                int32_t device) {
    if (!prm || P <= 0 || !starts_goals_host || !seeds_host || (prm->n_obst > 0 && !obst_host)) return RRTFND_E_BADARG;
    if (prm->stream_mode == RRTFND_STREAM_WORDS && !words_host) return RRTFND_E_BADARG;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) return RRTFND_E_NOGPU;
    CK(cudaSetDevice(device));
    const size_t C = (size_t)prm->capacity, M = (size_t)prm->n_obst, n = (size_t)P * C;

    std::vector<rrtfnd_planner_t> pl((size_t)P);
    memset(pl.data(), 0, pl.size() * sizeof(rrtfnd_planner_t));
    for (int p = 0; p < P; ++p) {
        pl[p].start_x = starts_goals_host[4 * p + 0]; pl[p].start_y = starts_goals_host[4 * p + 1];
        pl[p].goal_x = starts_goals_host[4 * p + 2];  pl[p].goal_y = starts_goals_host[4 * p + 3];
        pl[p].seed = seeds_host[2 * p]; pl[p].stream_id = seeds_host[2 * p + 1];
    }
    // obstacle rows ox, oy, r, K with K = ((-(r*r)) + ox*ox) + oy*oy (first three terms of `c`, src/algorithm.py:619)
    std::vector<double> ob(4 * M + 4);
    for (size_t j = 0; j < M; ++j) {
        const volatile double ox = obst_host[3 * j], oy = obst_host[3 * j + 1], r = obst_host[3 * j + 2];
        const volatile double r2 = r * r, x2 = ox * ox, y2 = oy * oy;
        const volatile double t = -r2 + x2;
        ob[4 * j] = ox; ob[4 * j + 1] = oy; ob[4 * j + 2] = r; ob[4 * j + 3] = t + y2;
    }

    DevBuf d_pl, d_x, d_y, d_e, d_par, d_nc, d_id, d_fin, d_bp, d_ob, d_w;
    CK(d_pl.alloc((size_t)P * sizeof(rrtfnd_planner_t)));
    CK(d_x.alloc(n * 8)); CK(d_y.alloc(n * 8)); CK(d_e.alloc(n * 8));
    CK(d_par.alloc(n * 4)); CK(d_nc.alloc(n * 4)); CK(d_id.alloc(n * 4));
    CK(d_fin.alloc((size_t)P * prm->finish_cap * 4)); CK(d_bp.alloc((size_t)P * prm->path_cap * 4));
    CK(d_ob.alloc(ob.size() * 8));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int rc = 0;
    do {
        if ((rc = (int)cudaMemcpyAsync(d_pl.p, pl.data(), (size_t)P * sizeof(rrtfnd_planner_t), cudaMemcpyHostToDevice, st))) break;
        if ((rc = (int)cudaMemcpyAsync(d_ob.p, ob.data(), ob.size() * 8, cudaMemcpyHostToDevice, st))) break;
        if (prm->stream_mode == RRTFND_STREAM_WORDS) {
            const size_t wb = (size_t)P * (size_t)prm->words_per_planner * 4;
            if ((rc = (int)d_w.alloc(wb))) break;
            if ((rc = (int)cudaMemcpyAsync(d_w.p, words_host, wb, cudaMemcpyHostToDevice, st))) break;
        }
        rrtfnd_batch_t bt;
        memset(&bt, 0, sizeof bt);
        bt.n_planners = P; bt.planners = (rrtfnd_planner_t*)d_pl.p;
        bt.x = (double*)d_x.p; bt.y = (double*)d_y.p; bt.ecost = (double*)d_e.p;
        bt.parent = (int32_t*)d_par.p; bt.nchild = (int32_t*)d_nc.p; bt.id = (int32_t*)d_id.p;
        bt.finish = (int32_t*)d_fin.p; bt.best_path = (int32_t*)d_bp.p; bt.obst = (const double*)d_ob.p;
        bt.words = (const uint32_t*)d_w.p; bt.trace = nullptr;
        rrtfnd_params_t q = *prm;
        q.flags &= ~RRTFND_FLAG_TRACE;
        if ((rc = rrtfnd_init_batch(&q, &bt, st))) break;
        if ((rc = rrtfnd_plan_batch(&q, &bt, st))) break;
        if (planners_host && (rc = (int)cudaMemcpyAsync(planners_host, d_pl.p, (size_t)P * sizeof(rrtfnd_planner_t), cudaMemcpyDeviceToHost, st))) break;
        if (x_host && (rc = (int)cudaMemcpyAsync(x_host, d_x.p, n * 8, cudaMemcpyDeviceToHost, st))) break;
        if (y_host && (rc = (int)cudaMemcpyAsync(y_host, d_y.p, n * 8, cudaMemcpyDeviceToHost, st))) break;
        if (ecost_host && (rc = (int)cudaMemcpyAsync(ecost_host, d_e.p, n * 8, cudaMemcpyDeviceToHost, st))) break;
        if (parent_host && (rc = (int)cudaMemcpyAsync(parent_host, d_par.p, n * 4, cudaMemcpyDeviceToHost, st))) break;
        if (nchild_host && (rc = (int)cudaMemcpyAsync(nchild_host, d_nc.p, n * 4, cudaMemcpyDeviceToHost, st))) break;
        if (id_host && (rc = (int)cudaMemcpyAsync(id_host, d_id.p, n * 4, cudaMemcpyDeviceToHost, st))) break;
        if (best_path_host && (rc = (int)cudaMemcpyAsync(best_path_host, d_bp.p, (size_t)P * prm->path_cap * 4, cudaMemcpyDeviceToHost, st))) break;
        rc = (int)cudaStreamSynchronize(st);
    } while (0);
    cudaStreamDestroy(st);
    return rc;
}
