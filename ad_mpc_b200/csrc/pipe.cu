// pipe.cu -- the chunk pipeline behind the C ABI: ONE handle, K internal chunk handles (one CUDA stream each).
//
// admpc_pipe_solve_host is the batched replacement of AD3DOptimizer.run_optimization (ad_mpc/ad_3d_optimizer.py:420-465)
// for B vehicles with HOST buffers: the batch is cut into K contiguous chunks; the H2D copy of chunk c+1, the two solver
// kernels of chunk c and the D2H copy of chunk c-1 overlap because every chunk owns a stream.  One call, one host
// thread; a C caller reaches the same end-to-end rate as the Python wrapper (which is now a thin shell over this).
#include <stdlib.h>
#include <vector>

#include "common.cuh"

struct admpc_pipe {
    std::vector<admpc_batch *> parts;
    std::vector<int> lo, hi;
    int B = 0, N = 0;
};

extern "C" int admpc_pipe_free(admpc_pipe *p)
{
    if (!p) return ADMPC_E_ARG;
    for (admpc_batch *h : p->parts) if (h) admpc_batch_free(h);
    delete p;
    return 0;
}

extern "C" int admpc_pipe_create(const admpc_opts *opts, int B, int device, int chunks, admpc_pipe **out)
{
    if (!opts || !out || B <= 0 || chunks <= 0) { admpc_set_error("admpc_pipe_create", "bad argument"); return ADMPC_E_ARG; }
    if (chunks > B) chunks = B;
    admpc_pipe *p = new admpc_pipe();
    p->B = B; p->N = opts->N;
    const int base = B / chunks, rem = B % chunks;          // contiguous blocks that differ by at most one instance
    for (int c = 0; c < chunks; c++) {
        const int lo = c * base + (c < rem ? c : rem), n = base + (c < rem ? 1 : 0);
        admpc_batch *h = nullptr;
        const int r = admpc_batch_create(opts, n, device, &h);
        if (r) { admpc_pipe_free(p); return r; }
        p->parts.push_back(h); p->lo.push_back(lo); p->hi.push_back(lo + n);
        // earlier chunks at higher stream priority: the chunks then finish in order instead of all together at the end, and the
        // read-back of chunk c overlaps the kernels of chunk c + 1 (ADMPC_PIPE_FLAT=1: equal priorities, the first version)
        if (!getenv("ADMPC_PIPE_FLAT")) {
            const int rr = admpc_batch_set_stream_level(h, c);
            if (rr) { admpc_pipe_free(p); return rr; }
        }
    }
    *out = p;
    return 0;
}

extern "C" int admpc_pipe_chunks(const admpc_pipe *p) { return p ? (int)p->parts.size() : ADMPC_E_ARG; }
extern "C" admpc_batch *admpc_pipe_chunk(admpc_pipe *p, int c, int *lo, int *hi)
{
    if (!p || c < 0 || c >= (int)p->parts.size()) return nullptr;
    if (lo) *lo = p->lo[c];
    if (hi) *hi = p->hi[c];
    return p->parts[c];
}

extern "C" int admpc_pipe_wait(admpc_pipe *p)
{
    if (!p) return ADMPC_E_ARG;
    int rc = 0;
    for (admpc_batch *h : p->parts) { const int r = admpc_batch_wait(h); if (r && !rc) rc = r; }
    return rc;
}

extern "C" int admpc_pipe_set_gp(admpc_pipe *p, int nout, int M, int dz, const int *feat, const int *rows, const double *X,
                                 const double *alpha, const double *ell, const double *sigma_f, const double *y_mean, int trig)
{
    if (!p) return ADMPC_E_ARG;
    for (admpc_batch *h : p->parts) {
        const int r = admpc_batch_set_gp(h, nout, M, dz, feat, rows, X, alpha, ell, sigma_f, y_mean, trig);
        if (r) return r;
    }
    return 0;
}

// host arrays are instance-major, so a chunk's rows are a plain pointer offset
extern "C" int admpc_pipe_set_iterate(admpc_pipe *p, const double *x, const double *u)
{
    if (!p) return ADMPC_E_ARG;
    const size_t nx = (size_t)(p->N + 1) * 7, nu = (size_t)p->N * 2;
    for (size_t c = 0; c < p->parts.size(); c++) {
        const int r = admpc_batch_set_iterate(p->parts[c], x ? x + p->lo[c] * nx : nullptr, u ? u + p->lo[c] * nu : nullptr);
        if (r) return r;
    }
    return 0;
}

extern "C" int admpc_pipe_set_track(admpc_pipe *p, int L, const double *traj, int H, double traj_dt, int anchor)
{
    if (!p) return ADMPC_E_ARG;
    for (admpc_batch *h : p->parts) {
        int r = admpc_batch_set_track_anchor(h, 0);
        if (!r) r = admpc_batch_set_track(h, L, traj, H, traj_dt);
        if (!r) r = admpc_batch_set_track_anchor(h, anchor);
        if (r) return r;
    }
    return 0;
}

// enqueue every chunk (H2D -> prepare -> QP -> D2H on the chunk's stream); returns without waiting
extern "C" int admpc_pipe_solve_host_async(admpc_pipe *p, const double *x0, const double *yref, const double *p_scalar,
                                           double *u_out, double *x_out, int *status_out)
{
    if (!p) return ADMPC_E_ARG;
    const size_t nx = (size_t)(p->N + 1) * 7, nu = (size_t)p->N * 2, ny = (size_t)p->N * 9 + 7;
    for (size_t c = 0; c < p->parts.size(); c++) {
        const size_t lo = p->lo[c];
        const int r = admpc_batch_solve_host_async(p->parts[c], x0 ? x0 + lo * 7 : nullptr, yref ? yref + lo * ny : nullptr,
                                                   p_scalar ? p_scalar + lo : nullptr, u_out ? u_out + lo * nu : nullptr,
                                                   x_out ? x_out + lo * nx : nullptr, status_out ? status_out + lo : nullptr);
        if (r) return r;
    }
    return 0;
}
extern "C" int admpc_pipe_solve_host(admpc_pipe *p, const double *x0, const double *yref, const double *p_scalar, double *u_out,
                                     double *x_out, int *status_out)
{
    const int r = admpc_pipe_solve_host_async(p, x0, yref, p_scalar, u_out, x_out, status_out);
    return r ? r : admpc_pipe_wait(p);
}

// pose-only control step per chunk: H2D(x0, p) -> reference generation -> solve -> D2H
extern "C" int admpc_pipe_solve_pose(admpc_pipe *p, const double *x0, const double *p_scalar, double *u_out, double *x_out,
                                     int *status_out)
{
    if (!p || !x0) return ADMPC_E_ARG;
    const size_t nx = (size_t)(p->N + 1) * 7, nu = (size_t)p->N * 2;
    for (size_t c = 0; c < p->parts.size(); c++) {
        const size_t lo = p->lo[c];
        const int r = admpc_batch_solve_pose_async(p->parts[c], x0 + lo * 7, p_scalar ? p_scalar + lo : nullptr,
                                                   u_out ? u_out + lo * nu : nullptr, x_out ? x_out + lo * nx : nullptr,
                                                   status_out ? status_out + lo : nullptr);
        if (r) return r;
    }
    return admpc_pipe_wait(p);
}

extern "C" long long admpc_pipe_kernel_launches(const admpc_pipe *p)
{
    long long n = 0;
    if (p) for (admpc_batch *h : p->parts) n += admpc_batch_kernel_launches(h);
    return n;
}
