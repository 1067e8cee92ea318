// api.cu -- C ABI of libadmpc_b200.so (include/admpc.h): batched handle + the acados-shim twin.
// Host side only orchestrates: device memory, one stream per handle, layout kernels, the three solver kernels.
// No CPU fallback exists: without a usable CUDA device every entry point returns ADMPC_E_CUDA.
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "common.cuh"

static thread_local char g_err[512] = "";
void admpc_set_error(const char *what, const char *msg) { snprintf(g_err, sizeof g_err, "%s: %s", what, msg); }
extern "C" const char *admpc_last_error(void) { return g_err; }

extern "C" int admpc_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" void admpc_default_opts(admpc_opts *o)
{
    memset(o, 0, sizeof *o);
    o->N = 20;
    o->iter_max = 50;
    o->dt = 0.05;
    const double W[9] = {10, 10, 100, 0, 0, 0, 0, 1, 100};     // acados_solver_sim_car.c:393-399
    const double We[7] = {1e-5, 1e-5, 1e-4, 0, 0, 0, 0};       // acados_solver_sim_car.c:481-485
    memcpy(o->W, W, sizeof W);
    memcpy(o->We, We, sizeof We);
    for (int j = 0; j < 2; j++) { o->zl[j] = o->zu[j] = 10.0; o->Zl[j] = o->Zu[j] = 0.0; }
    o->lbu[0] = -10; o->ubu[0] = 5; o->lbu[1] = -3; o->ubu[1] = 3;
    o->lbx = -0.52; o->ubx = 0.52;
    o->con_set = 0; o->lbx2 = -2.0; o->ubx2 = 2.0;
    // ad_3d.py:47-60 evaluated literally (the reference's pi literal is 3.14195)
    const double mass = 1500, f_mass = 900, r_mass = mass - f_mass, L = 2.7;
    o->mass = mass;
    o->lf = L * (1 - f_mass / mass);
    o->lr = L * (1 - r_mass / mass);
    o->iz = o->lf * o->lr * (r_mass + f_mass);
    o->cf2 = 2 * (f_mass * 0.5 * 9.81 * 0.165 * 180 / 3.14195);
    o->cr2 = 2 * (r_mass * 0.5 * 9.81 * 0.165 * 180 / 3.14195);
    o->mu0 = 10.0;
    o->tol_stat = o->tol_eq = o->tol_ineq = o->tol_comp = 1e-8;
    o->alpha_min = 1e-12;
    o->lam_min = o->t_min = 1e-16;
    o->thr0 = 0.1;
    o->reg = 1e-15;
    o->gp_feat[0] = 3; o->gp_feat[1] = 4; o->gp_feat[2] = 5; o->gp_feat[3] = 6;
    o->gp_row[0] = 4; o->gp_row[1] = 5;
}

// ---------------------------------------------------------------------------------------------- NCCL (dlopen) ---
typedef struct { char internal[128]; } nccl_uid;
typedef void *nccl_comm;
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(nccl_uid *) = nullptr;
    int (*CommInitRank)(nccl_comm *, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(nccl_comm) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load()
{
    if (g_nccl.lib) return 0;
    const char *cands[] = {getenv("ADMPC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *l = nullptr;
    for (const char *c : cands) {
        if (!c) continue;
        l = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (l) break;
    }
    if (!l) { admpc_set_error("dlopen(libnccl)", dlerror()); return ADMPC_E_NCCL; }
#define LOAD(name)                                                                                 \
    *(void **)(&g_nccl.name) = dlsym(l, "nccl" #name);                                             \
    if (!g_nccl.name) { admpc_set_error("dlsym", "nccl" #name); return ADMPC_E_NCCL; }
    LOAD(GetUniqueId) LOAD(CommInitRank) LOAD(CommDestroy) LOAD(Broadcast) LOAD(AllReduce) LOAD(Send) LOAD(Recv)
    LOAD(GroupStart) LOAD(GroupEnd) LOAD(GetErrorString)
#undef LOAD
    g_nccl.lib = l;
    return 0;
}
#define NCCL_CHECK_RET(call)                                                                              \
    do {                                                                                                  \
        int r_ = (call);                                                                                  \
        if (r_ != 0) { admpc_set_error(#call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?"); return ADMPC_E_NCCL; } \
    } while (0)
enum { NCCL_INT8 = 0, NCCL_INT32 = 2, NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MAX = 2 };

// ---------------------------------------------------------------------------------------------- batch handle ----
struct admpc_batch {
    Params P;
    int device = 0;
    cudaStream_t stream = nullptr;
    double *pool = nullptr;      // one allocation for every double array
    int *ipool = nullptr;
    double *stage_in = nullptr, *stage_u = nullptr, *stage_x = nullptr, *stage_misc = nullptr;
    int *stage_status = nullptr;
    int *gp_sel = nullptr;           // [Bp] cluster model per instance (GP ensemble)
    int *sqp_active = nullptr;       // device counters (one per SQP iteration parity), see admpc_batch_solve_sqp
    int *sqp_active_host = nullptr;  // pinned
    double *gp_blob = nullptr;
    size_t gp_blob_cap = 0;
    double *kap_sp = nullptr;        // Frenet variant: spline curvature table
    size_t kap_sp_cap = 0;
    double *gp_res = nullptr;        // per-stage GP results handed from gp_sweep_kernel to prepare_kernel
    size_t gp_res_cap = 0;
    double *l2_scratch = nullptr;
    size_t l2_bytes = 0;
    double *loop_prev_u = nullptr;      // [2N prev_u | 2 u_apply | 7 x_next][Bp]
    int *loop_i = nullptr;              // [has_prev | safe_count | valid | cmd_ok][Bp]
    double *track = nullptr, *track_info = nullptr;
    size_t track_cap = 0;
    int track_L = 0, track_H = 0, track_stop = 0, track_anchor = 0;
    double track_dt = 0.0;
    cudaEvent_t ev[8] = {};
    cudaEvent_t tm0 = nullptr, tm1 = nullptr;
    bool profiling = false;
    bool gps_set = false;
    int qp_variant = 0;      // 0 auto, 1 thread-per-instance (qp_ipm.cu), 4 warp(s) per instance, register-resident (qp_warp.cu, N <= 63), 7 resident warp(s) + DMMA sweeps (qp_mma.cu, N <= 127)
    long long launches = 0;
    float ms_solve = 0, ms_prepare = 0, ms_qp = 0;
    char *pack = nullptr, *gpack = nullptr;     // packed [u | x | status] block of this rank / of all ranks (root)
    // fused gather (admpc_batch_gather_enable): every rank writes its block straight into the root's gpack -- peer memory
    // opened through CUDA IPC -- from the epilogue of the QP kernel; admpc_batch_gather is then only the completion barrier
    char *gpack_peer = nullptr;                 // non-root: the root's gpack mapped into this process
    bool gat_on = false, gat_fresh = false;     // fused path active / block written by the last feedback kernel
    int gat_root = -1;
    // The root's block is DOUBLE-BUFFERED: writes (kernel epilogue or explicit pack) go to half gat_par, every gather call
    // flips it, so a rank that runs ahead never overwrites the half the root is still copying out (the all-reduce of the
    // NEXT gather orders the write after that: it cannot complete before the root has passed its previous D2H copies).
    int gat_par = 0, gat_last = 0;
    bool b_uniform = true;                      // every rank of the communicator holds the same number of instances
    int *bar_buf = nullptr;                     // 4-byte all-reduce scratch of the completion barrier
    nccl_comm comm = nullptr;
    int rank = 0, nranks = 1;
};

static size_t rup(size_t v, size_t m) { return (v + m - 1) / m * m; }
// bytes of one rank's packed [u | x | status] block (16-byte multiple so that every rank's slice stays aligned)
static size_t gather_block_bytes(const Params &P)
{
    const size_t nu = (size_t)P.B * P.o.N * 2, nx = (size_t)P.B * (P.o.N + 1) * 7;
    return rup((nu + nx) * sizeof(double) + (size_t)P.B * sizeof(int), 16);
}
// doubles per GP output in the packed blob: M points x (dz + 2) + tail (dz inverse squared length scales, y_mean)
static size_t gp_stride(int M, int dz) { return rup((size_t)M * (dz + 2) + dz + 1, 2); }
// whole blob: nout outputs, then the 2^(j/GP_TAB) table of the device exp2 (model.cuh)
static size_t gp_blob_doubles(int nout, int M, int dz) { return gp_stride(M, dz) * nout + GP_TAB; }

extern "C" int admpc_batch_free(admpc_batch *h);
// inside create: a failed CUDA call releases what was allocated so far
#define CREATE_CK(call)                                                                       \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) { admpc_set_error(#call, cudaGetErrorString(e_)); admpc_batch_free(h); return ADMPC_E_CUDA; } \
    } while (0)
extern "C" int admpc_batch_create(const admpc_opts *opts, int B, int device, admpc_batch **out)
{
    if (!opts || !out || B <= 0 || opts->N < 2 || opts->N > ADMPC_NMAX) { admpc_set_error("admpc_batch_create", "bad argument"); return ADMPC_E_ARG; }
    if (opts->model_variant != 0 && opts->model_variant != 1) { admpc_set_error("admpc_batch_create", "model_variant must be 0 (Cartesian) or 1 (Frenet)"); return ADMPC_E_ARG; }
    if (opts->con_set != 0 && opts->con_set != 1) { admpc_set_error("admpc_batch_create", "con_set must be 0 or 1"); return ADMPC_E_ARG; }
    if (opts->con_set == 1 && opts->model_variant != 1) {
        admpc_set_error("admpc_batch_create", "con_set = 1 (the Frenet variant's constraint set) needs model_variant = 1");
        return ADMPC_E_UNSUPPORTED;
    }
    int ndev = 0;
    CUDA_CHECK_RET(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { admpc_set_error("admpc_batch_create", "no such CUDA device"); return ADMPC_E_CUDA; }
    CUDA_CHECK_RET(cudaSetDevice(device));
    admpc_batch *h = new admpc_batch();
    h->device = device;
    if (const char *v = getenv("ADMPC_QP_VARIANT")) h->qp_variant = atoi(v);
    Params &P = h->P;
    memset(&P, 0, sizeof P);
    P.o = *opts;
    P.o.gp_enabled = 0;
    const int N = opts->N;
    P.B = B;
    P.Bp = (int)rup((size_t)B, 32);
    const size_t Bp = P.Bp;
    CREATE_CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (auto &e : h->ev) CREATE_CK(cudaEventCreate(&e));
    CREATE_CK(cudaEventCreate(&h->tm0));
    CREATE_CK(cudaEventCreate(&h->tm1));

    // carve the pool
    struct Item { double **p; size_t rows; };
    const size_t nX = (size_t)(N + 1) * 7, nU = (size_t)N * 2, nPi = (size_t)N * 7, nC = (size_t)N * con_rows(h->P.o);
    double *x0, *yref, *pp, *gps, *kap = nullptr;
    // Interface arrays always; the QP workspace only for handles that run the thread-per-instance kernel (a 35 KB/instance SoA
    // workspace streamed from HBM): the tensor-core kernels (variant 7, N <= 127) and the round-1 warp kernel (variant 4,
    // N <= 63) keep their whole working set on chip.
    int variant = h->qp_variant ? h->qp_variant : 7;
    if (variant != 7 && variant != 4) variant = 1;
    if ((variant == 7 && N > 127) || (variant == 4 && N > 63)) variant = 1;
    const bool frenet = h->P.o.model_variant == 1;
    const bool need_ws1 = (variant == 1) || frenet;
    // the shared-memory-resident tensor-core kernel (variant 7) stages instance-major linearisation records by TMA; every other
    // feedback kernel reads the SoA rows.  Exactly one of the two layouts exists per handle.
    const bool use_im = (variant == 7) && !frenet;
    // Frenet variant: the tensor-core kernel (qp_mma_g.cu) pulls 80-double instance-major records; written next to the dense
    // SoA linearisation the SQP residual kernel and the dense QP kernel read
    const bool use_im_f = frenet && (h->qp_variant == 0 || h->qp_variant == 7) && N <= 127;
    std::vector<Item> items = {
        {&x0, 7}, {&yref, (size_t)N * 9 + 7}, {&pp, (size_t)N}, {&gps, 7},
        {&P.xb, nX}, {&P.ub, nU}, {&P.pib, nPi}, {&P.lamb, nC}, {&P.tb, nC}, {&P.slb, nU}, {&P.sub, nU},
        {&P.nlp_res, 4},
        {&P.lin_d, frenet ? (size_t)(N + 1) * DL_ROWS : 0}, {&kap, frenet ? (size_t)N : 0},
        {&P.lin, (use_im || frenet) ? 0 : (size_t)(N + 1) * LIN_ROWS}, {&P.lin_im, use_im ? (size_t)(N + 1) * LIM_STRIDE : (use_im_f ? (size_t)(N + 1) * 80 : 0)}, {&P.res_out, 4},
    };
    if (need_ws1) {
        std::vector<Item> w1 = {
            {&P.dx, nX}, {&P.du, nU}, {&P.pi, nPi}, {&P.lam, nC}, {&P.t, nC}, {&P.sl, nU}, {&P.su, nU},
            {&P.rgu, nU}, {&P.rgx, nX}, {&P.rgsl, nU}, {&P.rgsu, nU}, {&P.rb, nPi}, {&P.rd, nC}, {&P.rm, nC},
            {&P.K, (size_t)N * 14}, {&P.Ginv, (size_t)N * 3}, {&P.P, (size_t)(N + 1) * 28}, {&P.Pb, nPi}, {&P.kf, nU}, {&P.pv, nX},
            {&P.ddu, nU}, {&P.ddx, nX}, {&P.dpi, nPi}, {&P.dlam, nC}, {&P.dt, nC}, {&P.dsl, nU}, {&P.dsu, nU},
        };
        items.insert(items.end(), w1.begin(), w1.end());
    }
    size_t rows = 0;
    for (auto &it : items) rows += it.rows;
    CREATE_CK(cudaMalloc(&h->pool, rows * Bp * sizeof(double)));
    CREATE_CK(cudaMemsetAsync(h->pool, 0, rows * Bp * sizeof(double), h->stream));
    size_t off = 0;
    for (auto &it : items) { *it.p = it.rows ? h->pool + off * Bp : nullptr; off += it.rows; }
    P.x0 = x0; P.yref = yref; P.p = pp; P.gps = gps; P.kappa = kap;
    CREATE_CK(cudaMalloc(&h->ipool, (7 * Bp + 32) * sizeof(int)));
    CREATE_CK(cudaMemsetAsync(h->ipool, 0, (7 * Bp + 32) * sizeof(int), h->stream));
    P.status = h->ipool; P.qp_status = h->ipool + Bp; P.qp_iter = h->ipool + 2 * Bp; P.lin_bad = h->ipool + 3 * Bp;
    P.sqp_status = h->ipool + 4 * Bp; P.sqp_iter = h->ipool + 5 * Bp;
    h->gp_sel = h->ipool + 6 * Bp; P.gp_sel = h->gp_sel;
    h->sqp_active = h->ipool + 7 * Bp;                       // per-iteration counters of still-running instances
    // instance-major staging areas
    const size_t in_rows = (size_t)N * 49 > (size_t)N * 9 + 7 ? (size_t)N * 49 : (size_t)N * 9 + 7;
    CREATE_CK(cudaMalloc(&h->stage_in, in_rows * Bp * sizeof(double)));
    CREATE_CK(cudaMalloc(&h->stage_u, nU * Bp * sizeof(double)));
    CREATE_CK(cudaMalloc(&h->stage_x, nX * Bp * sizeof(double)));
    CREATE_CK(cudaMalloc(&h->stage_misc, (size_t)N * 49 * Bp * sizeof(double)));
    CREATE_CK(cudaMalloc(&h->stage_status, 3 * Bp * sizeof(int)));
    CREATE_CK(cudaStreamSynchronize(h->stream));
    *out = h;
    return 0;
}
#undef CREATE_CK

extern "C" int admpc_batch_free(admpc_batch *h)
{
    if (!h) return ADMPC_E_ARG;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->gpack_peer) cudaIpcCloseMemHandle(h->gpack_peer);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    cudaFree(h->pool); cudaFree(h->ipool); cudaFree(h->stage_in); cudaFree(h->stage_u); cudaFree(h->stage_x);
    cudaFreeHost(h->sqp_active_host); cudaFree(h->stage_misc); cudaFree(h->stage_status); cudaFree(h->gp_blob); cudaFree(h->kap_sp); cudaFree(h->gp_res); cudaFree(h->l2_scratch); cudaFree(h->track); cudaFree(h->track_info); cudaFree(h->loop_prev_u); cudaFree(h->loop_i); cudaFree(h->pack); cudaFree(h->gpack); cudaFree(h->bar_buf);
    for (auto &e : h->ev) if (e) cudaEventDestroy(e);
    if (h->tm0) cudaEventDestroy(h->tm0);
    if (h->tm1) cudaEventDestroy(h->tm1);
    if (h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();          // a partially created handle may have produced benign errors above: do not leave them behind
    delete h;
    return 0;
}

extern "C" int admpc_batch_size(const admpc_batch *h) { return h ? h->P.B : ADMPC_E_ARG; }
extern "C" int admpc_batch_horizon(const admpc_batch *h) { return h ? h->P.o.N : ADMPC_E_ARG; }
extern "C" long long admpc_batch_kernel_launches(const admpc_batch *h) { return h ? h->launches : 0; }

// packs K cluster models (K = 1: a plain GP) into the TMA-friendly blob (see GpDev) and uploads it.
// Array layouts carry a leading model axis: X[K][nout][M][dz], alpha[K][nout][M], ell[K][nout][dz], sigma_f[K][nout],
// y_mean[K][nout], centroids[K][dz] (NULL for K = 1).
static int upload_gp(admpc_batch *h, int K, int nout, int M, int dz, const int *feat, const int *rows, const double *X,
                     const double *alpha, const double *ell, const double *sigma_f, const double *y_mean,
                     const double *centroids, int trig)
{
    Params &P = h->P;
    if (nout == 0) { P.o.gp_enabled = 0; return 0; }
    if (K < 1 || nout < 0 || nout > ADMPC_GPOUT_MAX || M <= 0 || dz <= 0 || dz > ADMPC_DZMAX || !feat || !rows || !X || !alpha || !ell || !sigma_f || !y_mean || (K > 1 && !centroids)) {
        admpc_set_error("admpc_batch_set_gp", "bad argument");
        return ADMPC_E_ARG;
    }
    for (int d = 0; d < dz; d++)
        if (feat[d] < 2 || feat[d] > 8) { admpc_set_error("admpc_batch_set_gp", "GP features must be in [psi..delta,u0,u1] (indices 2..8)"); return ADMPC_E_UNSUPPORTED; }
    for (int j = 0; j < nout; j++)
        if (rows[j] < 3 || rows[j] > 5) { admpc_set_error("admpc_batch_set_gp", "GP outputs must map to state rows 3..5"); return ADMPC_E_UNSUPPORTED; }
    const size_t stride = gp_stride(M, dz);
    const size_t model_doubles = stride * nout;
    const size_t total = model_doubles * K + GP_TAB;
    const size_t bytes = total * sizeof(double);
    if (bytes > 220 * 1024) { admpc_set_error("admpc_batch_set_gp", "GP model (all cluster models together) exceeds the shared-memory staging budget (220 KB)"); return ADMPC_E_UNSUPPORTED; }
    std::vector<double> blob(total, 0.0);
    for (int j = 0; j < GP_TAB; j++) blob[model_doubles * K + j] = exp2((double)j / GP_TAB);
    const double L2E = 1.4426950408889634;
    for (int c = 0; c < K; c++)
        for (int j = 0; j < nout; j++) {
            const size_t cj = (size_t)c * nout + j;
            double *b = blob.data() + model_doubles * c + stride * j;
            for (int i = 0; i < M; i++) {
                double cs = 0.0;
                for (int d = 0; d < dz; d++) {
                    const double x = X[(cj * M + i) * dz + d], wd = 1.0 / (ell[cj * dz + d] * ell[cj * dz + d]);
                    b[(size_t)i * (dz + 2) + d] = L2E * wd * x;
                    cs += wd * x * x;
                }
                b[(size_t)i * (dz + 2) + dz] = -0.5 * L2E * cs;
                b[(size_t)i * (dz + 2) + dz + 1] = sigma_f[cj] * alpha[cj * M + i];
            }
            double *w = b + (size_t)M * (dz + 2);
            for (int d = 0; d < dz; d++) w[d] = 1.0 / (ell[cj * dz + d] * ell[cj * dz + d]);
            w[dz] = y_mean[cj];
        }
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    const size_t cbytes = (size_t)K * dz * sizeof(double);
    if (bytes + cbytes > h->gp_blob_cap) {
        cudaFree(h->gp_blob);
        h->gp_blob = nullptr; h->gp_blob_cap = 0;                   // nothing may keep pointing at the freed block
        P.gp.blob = nullptr; P.o.gp_enabled = 0;
        CUDA_CHECK_RET(cudaMalloc(&h->gp_blob, bytes + cbytes));
        h->gp_blob_cap = bytes + cbytes;
    }
    CUDA_CHECK_RET(cudaMemcpyAsync(h->gp_blob, blob.data(), bytes, cudaMemcpyHostToDevice, h->stream));
    if (K > 1) CUDA_CHECK_RET(cudaMemcpyAsync((char *)h->gp_blob + bytes, centroids, cbytes, cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK_RET(cudaMemsetAsync(h->gp_sel, 0, (size_t)P.Bp * sizeof(int), h->stream));      // model 0 until a selection is made
    CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    P.gp.blob = h->gp_blob;
    P.gp.bytes = (int)bytes;
    P.gp.stride_out = (int)stride;
    P.gp.n_models = K;
    P.gp.model_doubles = (int)model_doubles;
    P.gp.centroids = (const double *)((char *)h->gp_blob + bytes);
    P.o.gp_enabled = 1;
    P.o.gp_nout = nout; P.o.gp_M = M; P.o.gp_dz = dz; P.o.gp_stage0_trigger = trig;
    for (int d = 0; d < dz; d++) P.o.gp_feat[d] = feat[d];
    for (int j = 0; j < nout; j++) P.o.gp_row[j] = rows[j];
    return 0;
}

// (re)allocates the stage-result buffer of the two-pass GP preparation for the model now configured in P.o
static int gp_res_alloc(admpc_batch *h)
{
    Params &P = h->P;
    if (!P.o.gp_enabled) return 0;
    const size_t need = (size_t)P.o.N * 4 * P.o.gp_nout * (1 + P.o.gp_dz) * P.Bp * sizeof(double);
    if (need > h->gp_res_cap) {
        cudaFree(h->gp_res);
        h->gp_res = nullptr; h->gp_res_cap = 0; P.gpr = nullptr;
        if (cudaMalloc(&h->gp_res, need) != cudaSuccess) {
            cudaGetLastError();
            P.o.gp_enabled = 0;
            admpc_set_error("GP stage-result buffer", "cudaMalloc failed");
            return ADMPC_E_CUDA;
        }
        h->gp_res_cap = need;
    }
    P.gpr = h->gp_res;
    return 0;
}

extern "C" int admpc_batch_set_gp(admpc_batch *h, int nout, int M, int dz, const int *feat, const int *rows,
                                  const double *X, const double *alpha, const double *ell, const double *sigma_f,
                                  const double *y_mean, int stage0_trigger)
{
    if (!h) return ADMPC_E_ARG;
    const int r = upload_gp(h, 1, nout, M, dz, feat, rows, X, alpha, ell, sigma_f, y_mean, nullptr, stage0_trigger);
    return r ? r : gp_res_alloc(h);
}

// GP ensemble (GPEnsemble, gp.py:536-770; homogeneous case: the same K clusters for every output dimension):
// K cluster models with their centroids in feature space; every instance uses ONE of them per solve, chosen by
// admpc_batch_select_gp (nearest centroid, gp.py:738-770) or set explicitly (the reference's use_model argument).
extern "C" int admpc_batch_set_gp_ensemble(admpc_batch *h, int K, int nout, int M, int dz, const int *feat, const int *rows,
                                           const double *X, const double *alpha, const double *ell, const double *sigma_f,
                                           const double *y_mean, const double *centroids, int stage0_trigger)
{
    if (!h) return ADMPC_E_ARG;
    const int r = upload_gp(h, K, nout, M, dz, feat, rows, X, alpha, ell, sigma_f, y_mean, centroids, stage0_trigger);
    return r ? r : gp_res_alloc(h);
}

// nearest-centroid choice per instance from the query state xq [B][7] (NULL: the current x0) and input uq [B][2]
// (NULL: zeros): argmin_c || z - centroid_c ||, z = B_z [x; u]; first minimum wins like np.argmin.
extern "C" int admpc_batch_select_gp(admpc_batch *h, const double *xq, const double *uq)
{
    if (!h) return ADMPC_E_ARG;
    Params &P = h->P;
    if (!P.o.gp_enabled) { admpc_set_error("admpc_batch_select_gp", "no GP model set"); return ADMPC_E_STATE; }
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    const double *dx = P.x0, *du = nullptr;
    if (xq) {
        CUDA_CHECK_RET(cudaMemcpyAsync(h->stage_in, xq, (size_t)P.B * 7 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        launch_transpose_in(h->stage_in, h->stage_x, P.B, P.Bp, 7, h->stream);
        dx = h->stage_x;
    }
    if (uq) {
        CUDA_CHECK_RET(cudaMemcpyAsync(h->stage_misc, uq, (size_t)P.B * 2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        launch_transpose_in(h->stage_misc, h->stage_u, P.B, P.Bp, 2, h->stream);
        du = h->stage_u;
    }
    launch_gp_select(P, dx, du, h->gp_sel, h->stream);
    h->launches += 1 + (xq ? 1 : 0) + (uq ? 1 : 0);
    CUDA_CHECK_RET(cudaGetLastError());
    return 0;
}
extern "C" int admpc_batch_set_gp_index(admpc_batch *h, const int *idx)
{
    if (!h || !idx) return ADMPC_E_ARG;
    const Params &P = h->P;
    for (int i = 0; i < P.B; i++)
        if (idx[i] < 0 || idx[i] >= (P.gp.n_models > 0 ? P.gp.n_models : 1)) { admpc_set_error("admpc_batch_set_gp_index", "model index out of range"); return ADMPC_E_ARG; }
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    CUDA_CHECK_RET(cudaMemcpyAsync(h->gp_sel, idx, (size_t)P.B * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    return 0;
}
extern "C" int admpc_batch_get_gp_index(admpc_batch *h, int *idx)
{
    if (!h || !idx) return ADMPC_E_ARG;
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    CUDA_CHECK_RET(cudaMemcpyAsync(idx, h->gp_sel, (size_t)h->P.B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    return 0;
}

// host instance-major [B][F] -> device SoA rows
static int put_rows(admpc_batch *h, const double *host, double *dst, int F)
{
    if (!h || !host) { admpc_set_error("admpc set", "null pointer"); return ADMPC_E_ARG; }
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    CUDA_CHECK_RET(cudaMemcpyAsync(h->stage_in, host, (size_t)h->P.B * F * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    launch_transpose_in(h->stage_in, dst, h->P.B, h->P.Bp, F, h->stream);
    h->launches++;
    CUDA_CHECK_RET(cudaGetLastError());
    return 0;
}

extern "C" int admpc_batch_set_x0(admpc_batch *h, const double *x0)
{
    if (!h) return ADMPC_E_ARG;
    int r = put_rows(h, x0, (double *)h->P.x0, 7);
    if (r == 0 && !h->gps_set) {   // gp_state defaults to the initial state (quad_3d_optimizer.py:549)
        CUDA_CHECK_RET(cudaMemcpyAsync((void *)h->P.gps, h->P.x0, (size_t)7 * h->P.Bp * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    }
    return r;
}
extern "C" int admpc_batch_set_yref(admpc_batch *h, const double *yref) { return h ? put_rows(h, yref, (double *)h->P.yref, h->P.o.N * 9 + 7) : ADMPC_E_ARG; }
extern "C" int admpc_batch_set_p(admpc_batch *h, const double *p) { return h ? put_rows(h, p, (double *)h->P.p, h->P.o.N) : ADMPC_E_ARG; }
extern "C" int admpc_batch_set_kappa(admpc_batch *h, const double *kappa)
{
    if (!h) return ADMPC_E_ARG;
    if (h->P.o.model_variant != 1) { admpc_set_error("admpc_batch_set_kappa", "handle was not created with model_variant = 1 (Frenet)"); return ADMPC_E_STATE; }
    return put_rows(h, kappa, (double *)h->P.kappa, h->P.o.N);
}
// Frenet variant: kappa(s) as a per-instance piecewise cubic (K pieces: breaks[B][K+1], coef[B][K][4], lowest power first),
// evaluated inside the model at every RK4 sub-stage with its d kappa / d s Jacobian column -- the reference's
// interpolant('kapparef_s', 'bspline', ...) semantics (fren_ad_3d_optimizer bytecode) for any spline brought to
// piecewise-polynomial form.  K = 0 / NULL switches back to the per-node constants of admpc_batch_set_kappa.
extern "C" int admpc_batch_set_kappa_spline(admpc_batch *h, int K, const double *breaks, const double *coef)
{
    if (!h) return ADMPC_E_ARG;
    Params &P = h->P;
    if (P.o.model_variant != 1) { admpc_set_error("admpc_batch_set_kappa_spline", "handle was not created with model_variant = 1 (Frenet)"); return ADMPC_E_STATE; }
    if (K <= 0 || !breaks || !coef) { P.kap_K = 0; return 0; }
    if (K > 64) { admpc_set_error("admpc_batch_set_kappa_spline", "at most 64 pieces"); return ADMPC_E_UNSUPPORTED; }
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    const size_t rows = (size_t)(K + 1) + 4 * (size_t)K, need = rows * P.Bp * sizeof(double);
    if (need > h->kap_sp_cap) {
        cudaFree(h->kap_sp);
        h->kap_sp = nullptr; h->kap_sp_cap = 0; P.kap_K = 0; P.kap_sp = nullptr;
        CUDA_CHECK_RET(cudaMalloc(&h->kap_sp, need));
        h->kap_sp_cap = need;
    }
    // host [B][K+1] and [B][K*4] -> SoA rows, through the staging area in two transposes
    double *tmp = nullptr;
    CUDA_CHECK_RET(cudaMalloc(&tmp, (size_t)P.B * (4 * K) * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(tmp, breaks, (size_t)P.B * (K + 1) * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) { launch_transpose_in(tmp, h->kap_sp, P.B, P.Bp, K + 1, h->stream); e = cudaStreamSynchronize(h->stream); }
    if (e == cudaSuccess) e = cudaMemcpyAsync(tmp, coef, (size_t)P.B * 4 * K * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) { launch_transpose_in(tmp, h->kap_sp + (size_t)(K + 1) * P.Bp, P.B, P.Bp, 4 * K, h->stream); e = cudaStreamSynchronize(h->stream); }
    cudaFree(tmp);
    h->launches += 2;
    if (e != cudaSuccess) { admpc_set_error("admpc_batch_set_kappa_spline", cudaGetErrorString(e)); return ADMPC_E_CUDA; }
    P.kap_sp = h->kap_sp; P.kap_K = K;
    return 0;
}
extern "C" int admpc_batch_set_p_scalar(admpc_batch *h, const double *p)
{
    if (!h || !p) return ADMPC_E_ARG;
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    CUDA_CHECK_RET(cudaMemsetAsync(h->stage_in, 0, (size_t)h->P.Bp * sizeof(double), h->stream));
    CUDA_CHECK_RET(cudaMemcpyAsync(h->stage_in, p, (size_t)h->P.B * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    launch_bcast_rows(h->stage_in, (double *)h->P.p, h->P.Bp, h->P.o.N, h->stream);
    h->launches++;
    CUDA_CHECK_RET(cudaGetLastError());
    return 0;
}
extern "C" int admpc_batch_set_gp_state(admpc_batch *h, const double *gp_state)
{
    if (!h) return ADMPC_E_ARG;
    if (!gp_state) { h->gps_set = false; return 0; }
    h->gps_set = true;
    return put_rows(h, gp_state, (double *)h->P.gps, 7);
}
extern "C" int admpc_batch_set_iterate(admpc_batch *h, const double *x, const double *u)
{
    if (!h) return ADMPC_E_ARG;
    int r = 0;
    h->gat_fresh = false;
    if (x) r = put_rows(h, x, h->P.xb, (h->P.o.N + 1) * 7);
    if (r == 0 && u) r = put_rows(h, u, h->P.ub, h->P.o.N * 2);
    return r;
}
// multipliers and slacks of the iterate (acados load_iterate restores all of x, u, pi, lam, t, sl, su:
// ad_3d_optimizer.py:454).  The RTI step overwrites them with the QP's; the full-SQP loop's first KKT check reads them.
// Host layouts [B][N*7], [B][N*10], [B][N*10], [B][N*2], [B][N*2]; any pointer may be NULL (left unchanged).
extern "C" int admpc_batch_set_duals(admpc_batch *h, const double *pi, const double *lam, const double *t, const double *sl,
                                     const double *su)
{
    if (!h) return ADMPC_E_ARG;
    const int N = h->P.o.N;
    int r = 0;
    if (pi) r = put_rows(h, pi, h->P.pib, N * 7);
    if (r == 0 && lam) r = put_rows(h, lam, h->P.lamb, N * con_rows(h->P.o));
    if (r == 0 && t) r = put_rows(h, t, h->P.tb, N * con_rows(h->P.o));
    if (r == 0 && sl) r = put_rows(h, sl, h->P.slb, N * 2);
    if (r == 0 && su) r = put_rows(h, su, h->P.sub, N * 2);
    return r;
}
// internal (pipe.cu): give the handle's stream a scheduling priority (0 = highest level the device offers, larger = lower; clamped
// to the device's range).  The chunk pipeline runs chunk c at a higher priority than chunk c + 1, so that the CTA scheduler
// finishes the chunks in order and every read-back overlaps the next chunk's kernels.
int admpc_batch_set_stream_level(admpc_batch *h, int level)
{
    if (!h) return ADMPC_E_ARG;
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    int least = 0, greatest = 0;
    CUDA_CHECK_RET(cudaDeviceGetStreamPriorityRange(&least, &greatest));      // numerically: greatest <= least
    int prio = greatest + level;
    if (prio > least) prio = least;
    cudaStream_t ns = nullptr;
    CUDA_CHECK_RET(cudaStreamCreateWithPriority(&ns, cudaStreamNonBlocking, prio));
    CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    cudaStreamDestroy(h->stream);
    h->stream = ns;
    return 0;
}

extern "C" int admpc_batch_reset(admpc_batch *h)
{
    if (!h) return ADMPC_E_ARG;
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    const Params &P = h->P;
    // iterate arrays are contiguous in the pool: xb .. sub
    const size_t n = (size_t)((P.sub + (size_t)P.o.N * 2 * P.Bp) - P.xb);
    h->gat_fresh = false;
    CUDA_CHECK_RET(cudaMemsetAsync(P.xb, 0, n * sizeof(double), h->stream));
    return 0;
}

// feedback phase + update of one iteration (shared by the RTI step and the full-SQP loop)
static int launch_feedback(admpc_batch *h)
{
    const Params &P = h->P;
    h->gat_fresh = false;
    if (P.o.model_variant == 1) {
        // Frenet variant: the tensor-core kernel qp_mma_g (N <= 127, no trivial column assumed -- a spline curvature makes the
        // column of s dense --, both constraint sets, fused update); on request the round-1 warp kernel on the 6x8 stage
        // structure (ADMPC_QP_VARIANT=4: per-node curvature and con_set = 0 only) or the dense thread-per-instance kernel +
        // separate update (ADMPC_QP_VARIANT=1, also the N > 127 fallback)
        if ((h->qp_variant == 0 || h->qp_variant == 7) && launch_qp_mma_g(P, h->stream)) {
            if (h->profiling) CUDA_CHECK_RET(cudaEventRecord(h->ev[3], h->stream));
            h->launches += 1;
            h->gat_fresh = h->gat_on;
            return 0;
        }
        if (h->qp_variant != 1 && P.kap_K == 0 && P.o.con_set == 0 && launch_qp_warp_f(P, h->stream)) {
            if (h->profiling) CUDA_CHECK_RET(cudaEventRecord(h->ev[3], h->stream));
            h->launches += 1;
            h->gat_fresh = h->gat_on;
            return 0;
        }
        launch_qp_dense(P, h->stream);
        if (h->profiling) CUDA_CHECK_RET(cudaEventRecord(h->ev[3], h->stream));
        launch_update(P, h->stream);
        h->launches += 2;
        return 0;
    }
    // QP variant (0 = auto): 7 one / two / four warps per instance, shared-memory resident, Riccati sweeps as FP64 tensor-core
    // fragments (qp_mma.cu, default for N <= 127); 4 one / two warps per instance with register-resident IPM state (qp_warp.cu,
    // the round-1 kernel, kept as an independent implementation, N <= 63); 1 one thread per instance (any horizon).  A variant
    // that cannot take the horizon falls through to the thread-per-instance kernel.
    int variant = h->qp_variant ? h->qp_variant : 7;
    bool fused = false;
    if (variant == 7) { fused = launch_qp_mma(P, h->stream); h->gat_fresh = fused && h->gat_on; }    // sweeps on the FP64 tensor cores
    if (!fused && variant == 4) { fused = launch_qp_warp(P, h->stream); h->gat_fresh = fused && h->gat_on; }   // N <= 63
    if (!fused) {
        if (!P.dx) { admpc_set_error("admpc_batch_solve", "QP workspace for this kernel variant was not allocated at create"); return ADMPC_E_STATE; }
        launch_qp(P, h->stream);
    }
    if (h->profiling) CUDA_CHECK_RET(cudaEventRecord(h->ev[3], h->stream));
    if (!fused) launch_update(P, h->stream);
    h->launches += fused ? 1 : 2;
    return 0;
}

extern "C" int admpc_batch_solve(admpc_batch *h)
{
    if (!h) return ADMPC_E_ARG;
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    // Frenet variant, RTI step through the tensor-core kernel: only the instance-major records are consumed
    h->P.skip_lin_d = (h->P.o.model_variant == 1 && (h->qp_variant == 0 || h->qp_variant == 7) && h->P.o.N <= 127 && h->P.lin_im) ? 1 : 0;
    const Params &P = h->P;
    CUDA_CHECK_RET(cudaEventRecord(h->ev[0], h->stream));
    CUDA_CHECK_RET(cudaMemsetAsync(P.lin_bad, 0, (size_t)P.Bp * sizeof(int), h->stream));
    if (h->profiling) CUDA_CHECK_RET(cudaEventRecord(h->ev[1], h->stream));
    if (P.o.model_variant == 1) launch_prepare_dense(P, h->stream);
    else launch_prepare(P, h->stream);
    h->launches += P.o.gp_enabled ? 2 : 1;      // GP: sweep kernel + sensitivity kernel
    if (h->profiling) CUDA_CHECK_RET(cudaEventRecord(h->ev[2], h->stream));
    if (int r = launch_feedback(h)) return r;
    CUDA_CHECK_RET(cudaEventRecord(h->ev[4], h->stream));
    CUDA_CHECK_RET(cudaGetLastError());
    return 0;
}

// Full SQP (nlp_solver_type "SQP", create_ros_ad_mpc.py:47-51): repeat { prepare; NLP residual check; QP; full step }
// until every instance has converged / failed or max_iter is reached.  Synchronous: the host reads one counter per
// iteration (how many instances are still running).  Per-instance results: admpc_batch_get_sqp_info.
static int solve_sqp_impl(admpc_batch *h, int max_iter, const double *tol4, int *iterations_run)
{
    if (!h || max_iter < 0) return ADMPC_E_ARG;
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    h->P.skip_lin_d = 0;                                     // the NLP residual kernel reads the SoA linearisation
    const Params &P = h->P;
    const double dflt[4] = {1e-6, 1e-6, 1e-6, 1e-6};            // sim_car_acados_ocp.json:870-873
    const double *tol = tol4 ? tol4 : dflt;
    if (!h->sqp_active_host) CUDA_CHECK_RET(cudaHostAlloc((void **)&h->sqp_active_host, 32 * sizeof(int), cudaHostAllocDefault));
    CUDA_CHECK_RET(cudaEventRecord(h->ev[0], h->stream));
    CUDA_CHECK_RET(cudaMemsetAsync(P.lin_bad, 0, (size_t)P.Bp * sizeof(int), h->stream));
    CUDA_CHECK_RET(cudaMemsetAsync(P.status, 0, (size_t)P.Bp * sizeof(int), h->stream));
    CUDA_CHECK_RET(cudaMemsetAsync(P.sqp_iter, 0, (size_t)P.Bp * sizeof(int), h->stream));
    int it = 0, running = P.B;
    for (it = 0; it < max_iter; it++) {
        int *ctr = h->sqp_active + (it & 1);
        CUDA_CHECK_RET(cudaMemsetAsync(ctr, 0, sizeof(int), h->stream));
        if (P.o.model_variant == 1) { launch_prepare_dense(P, h->stream); launch_nlp_res_dense(P, it, tol, ctr, h->stream); }
        else { launch_prepare(P, h->stream); launch_nlp_res(P, it, tol, ctr, h->stream); }
        h->launches += P.o.gp_enabled ? 3 : 2;
        CUDA_CHECK_RET(cudaMemcpyAsync(h->sqp_active_host, ctr, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
        running = h->sqp_active_host[0];
        if (running == 0) break;
        if (int r = launch_feedback(h)) return r;
    }
    if (running != 0) {              // max_iter reached (or max_iter == 0): classify what is still running
        launch_sqp_finalize(P, h->stream);
        h->launches += 1;
    }
    CUDA_CHECK_RET(cudaEventRecord(h->ev[4], h->stream));
    CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    CUDA_CHECK_RET(cudaGetLastError());
    if (iterations_run) *iterations_run = it;
    return 0;
}
extern "C" int admpc_batch_solve_sqp(admpc_batch *h, int max_iter, const double *tol4, int *iterations_run)
{
    const int r = solve_sqp_impl(h, max_iter, tol4, iterations_run);
    if (h) h->gat_fresh = false;      // the iterate moved after the last fused-gather epilogue
    return r;
}

// per-instance outcome of the last admpc_batch_solve_sqp: acados status {0 converged, 1 NaN, 2 max_iter, 4 QP failure},
// number of QPs solved, NLP residual norms (stat, eq, ineq, comp) of the last check.  Any pointer may be NULL.
extern "C" int admpc_batch_get_sqp_info(admpc_batch *h, int *status, int *sqp_iter, double *res /*[B][4]*/)
{
    if (!h) return ADMPC_E_ARG;
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    const Params &P = h->P;
    CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    if (status) CUDA_CHECK_RET(cudaMemcpy(status, P.sqp_status, (size_t)P.B * sizeof(int), cudaMemcpyDeviceToHost));
    if (sqp_iter) CUDA_CHECK_RET(cudaMemcpy(sqp_iter, P.sqp_iter, (size_t)P.B * sizeof(int), cudaMemcpyDeviceToHost));
    if (res) {
        launch_transpose_out(P.nlp_res, h->stage_misc, P.B, P.Bp, 4, h->stream);
        CUDA_CHECK_RET(cudaMemcpyAsync(res, h->stage_misc, (size_t)P.B * 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    }
    return 0;
}

extern "C" int admpc_batch_wait(admpc_batch *h)
{
    if (!h) return ADMPC_E_ARG;
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    return 0;
}

static int get_rows_async(admpc_batch *h, const double *src, double *stage, double *host, int F)
{
    launch_transpose_out(src, stage, h->P.B, h->P.Bp, F, h->stream);
    h->launches++;
    CUDA_CHECK_RET(cudaGetLastError());
    CUDA_CHECK_RET(cudaMemcpyAsync(host, stage, (size_t)h->P.B * F * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    return 0;
}
static int get_rows(admpc_batch *h, const double *src, double *host, int F)
{
    if (!h || !host) return ADMPC_E_ARG;
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    int r = get_rows_async(h, src, h->stage_misc, host, F);
    if (r) return r;
    CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    return 0;
}
extern "C" int admpc_batch_get_u(admpc_batch *h, double *u) { return h ? get_rows(h, h->P.ub, u, h->P.o.N * 2) : ADMPC_E_ARG; }
extern "C" int admpc_batch_get_x(admpc_batch *h, double *x) { return h ? get_rows(h, h->P.xb, x, (h->P.o.N + 1) * 7) : ADMPC_E_ARG; }
extern "C" int admpc_batch_get_pi(admpc_batch *h, double *pi) { return h ? get_rows(h, h->P.pib, pi, h->P.o.N * 7) : ADMPC_E_ARG; }
extern "C" int admpc_batch_get_lam(admpc_batch *h, double *lam) { return h ? get_rows(h, h->P.lamb, lam, h->P.o.N * con_rows(h->P.o)) : ADMPC_E_ARG; }
extern "C" int admpc_batch_get_t(admpc_batch *h, double *t) { return h ? get_rows(h, h->P.tb, t, h->P.o.N * con_rows(h->P.o)) : ADMPC_E_ARG; }
extern "C" int admpc_batch_get_slacks(admpc_batch *h, double *sl, double *su)
{
    if (!h) return ADMPC_E_ARG;
    int r = 0;
    if (sl) r = get_rows(h, h->P.slb, sl, h->P.o.N * 2);
    if (r == 0 && su) r = get_rows(h, h->P.sub, su, h->P.o.N * 2);
    return r;
}
extern "C" int admpc_batch_get_status(admpc_batch *h, int *status, int *qp_status, int *qp_iter)
{
    if (!h) return ADMPC_E_ARG;
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    const size_t n = (size_t)h->P.B * sizeof(int);
    if (status) CUDA_CHECK_RET(cudaMemcpyAsync(status, h->P.status, n, cudaMemcpyDeviceToHost, h->stream));
    if (qp_status) CUDA_CHECK_RET(cudaMemcpyAsync(qp_status, h->P.qp_status, n, cudaMemcpyDeviceToHost, h->stream));
    if (qp_iter) CUDA_CHECK_RET(cudaMemcpyAsync(qp_iter, h->P.qp_iter, n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    return 0;
}

// expands the structured linearisation to dense row-major A[7x7], B[7x2] per stage (parity tests)
__global__ void expand_lin_kernel(const Params P, double *A, double *Bm, double *b, double *q, double *r)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    const int N = P.o.N;
    if (i >= P.B) return;
    if (k < N) {
        double *Ao = A + ((size_t)i * N + k) * 49, *Bo = Bm + ((size_t)i * N + k) * 14;
        for (int rr = 0; rr < 7; rr++)
            for (int c = 0; c < 7; c++) {
                double v;
                if (c < 2) v = (rr == c) ? 1.0 : 0.0;
                else if (rr == 6) v = (c == 6) ? 1.0 : 0.0;
                else v = lin_get(P, k, LIN_A + rr * 5 + (c - 2), i);
                Ao[rr * 7 + c] = v;
            }
        for (int rr = 0; rr < 7; rr++)
            for (int c = 0; c < 2; c++) Bo[rr * 2 + c] = (rr == 6) ? ((c == 1) ? P.o.dt : 0.0) : lin_get(P, k, LIN_B + rr * 2 + c, i);
        for (int c = 0; c < 7; c++) b[((size_t)i * N + k) * 7 + c] = lin_get(P, k, LIN_b + c, i);
        for (int c = 0; c < 2; c++) r[((size_t)i * N + k) * 2 + c] = lin_get(P, k, LIN_r + c, i);
    }
    for (int c = 0; c < 7; c++) q[((size_t)i * (N + 1) + k) * 7 + c] = lin_get(P, k, LIN_q + c, i);
}

extern "C" int admpc_batch_get_lin(admpc_batch *h, double *A, double *Bm, double *b, double *q, double *r)
{
    if (!h || !A || !Bm || !b || !q || !r) return ADMPC_E_ARG;
    if (h->P.o.model_variant != 0) { admpc_set_error("admpc_batch_get_lin", "structured linearisation of the Cartesian model only"); return ADMPC_E_UNSUPPORTED; }
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    const int N = h->P.o.N, B = h->P.B;
    const size_t nA = (size_t)B * N * 49, nB = (size_t)B * N * 14, nb = (size_t)B * N * 7, nq = (size_t)B * (N + 1) * 7, nr = (size_t)B * N * 2;
    double *d = nullptr;
    CUDA_CHECK_RET(cudaMalloc(&d, (nA + nB + nb + nq + nr) * sizeof(double)));
    dim3 grid((B + 127) / 128, N + 1);
    expand_lin_kernel<<<grid, 128, 0, h->stream>>>(h->P, d, d + nA, d + nA + nB, d + nA + nB + nb, d + nA + nB + nb + nq);
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(A, d, nA * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(Bm, d + nA, nB * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(b, d + nA + nB, nb * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(q, d + nA + nB + nb, nq * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(r, d + nA + nB + nb + nq, nr * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    if (e != cudaSuccess) { admpc_set_error("admpc_batch_get_lin", cudaGetErrorString(e)); return ADMPC_E_CUDA; }
    return 0;
}

static int solve_host_enqueue(admpc_batch *h, const double *x0, const double *yref, const double *p_scalar,
                              double *u_out, double *x_out, int *status_out);

extern "C" int admpc_batch_solve_host_async(admpc_batch *h, const double *x0, const double *yref, const double *p_scalar,
                                            double *u_out, double *x_out, int *status_out)
{
    return solve_host_enqueue(h, x0, yref, p_scalar, u_out, x_out, status_out);
}

extern "C" int admpc_batch_solve_host(admpc_batch *h, const double *x0, const double *yref, const double *p_scalar,
                                      double *u_out, double *x_out, int *status_out)
{
    int r = solve_host_enqueue(h, x0, yref, p_scalar, u_out, x_out, status_out);
    if (r) return r;
    CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    return 0;
}

static int solve_host_enqueue(admpc_batch *h, const double *x0, const double *yref, const double *p_scalar,
                              double *u_out, double *x_out, int *status_out)
{
    if (!h) return ADMPC_E_ARG;
    int r = 0;
    if (x0 && (r = admpc_batch_set_x0(h, x0))) return r;
    if (yref && (r = admpc_batch_set_yref(h, yref))) return r;
    if (p_scalar && (r = admpc_batch_set_p_scalar(h, p_scalar))) return r;
    if ((r = admpc_batch_solve(h))) return r;
    if (u_out && (r = get_rows_async(h, h->P.ub, h->stage_u, u_out, h->P.o.N * 2))) return r;
    if (x_out && (r = get_rows_async(h, h->P.xb, h->stage_x, x_out, (h->P.o.N + 1) * 7))) return r;
    if (status_out) CUDA_CHECK_RET(cudaMemcpyAsync(status_out, h->P.status, (size_t)h->P.B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    return 0;
}

extern "C" int admpc_batch_set_profiling(admpc_batch *h, int on) { if (!h) return ADMPC_E_ARG; h->profiling = on != 0; return 0; }

extern "C" int admpc_batch_last_ms(admpc_batch *h, const char *name, float *ms)
{
    if (!h || !name || !ms) return ADMPC_E_ARG;
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    CUDA_CHECK_RET(cudaEventSynchronize(h->ev[4]));
    if (!strcmp(name, "solve")) { CUDA_CHECK_RET(cudaEventElapsedTime(ms, h->ev[0], h->ev[4])); return 0; }
    if (!h->profiling) { admpc_set_error("admpc_batch_last_ms", "profiling is off"); return ADMPC_E_STATE; }
    if (!strcmp(name, "prepare")) { CUDA_CHECK_RET(cudaEventElapsedTime(ms, h->ev[1], h->ev[2])); return 0; }
    if (!strcmp(name, "qp")) { CUDA_CHECK_RET(cudaEventElapsedTime(ms, h->ev[2], h->ev[3])); return 0; }
    if (!strcmp(name, "update")) { CUDA_CHECK_RET(cudaEventElapsedTime(ms, h->ev[3], h->ev[4])); return 0; }
    admpc_set_error("admpc_batch_last_ms", "unknown timer");
    return ADMPC_E_ARG;
}

extern "C" int admpc_batch_timer_start(admpc_batch *h)
{
    if (!h) return ADMPC_E_ARG;
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    CUDA_CHECK_RET(cudaEventRecord(h->tm0, h->stream));
    return 0;
}
extern "C" int admpc_batch_timer_stop(admpc_batch *h, float *ms)
{
    if (!h || !ms) return ADMPC_E_ARG;
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    CUDA_CHECK_RET(cudaEventRecord(h->tm1, h->stream));
    CUDA_CHECK_RET(cudaEventSynchronize(h->tm1));
    CUDA_CHECK_RET(cudaEventElapsedTime(ms, h->tm0, h->tm1));
    return 0;
}
extern "C" int admpc_batch_flush_l2(admpc_batch *h)
{
    if (!h) return ADMPC_E_ARG;
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    if (!h->l2_scratch) {
        h->l2_bytes = (size_t)256 << 20;   // 256 MiB > 126 MB L2
        CUDA_CHECK_RET(cudaMalloc(&h->l2_scratch, h->l2_bytes));
    }
    launch_fill(h->l2_scratch, h->l2_bytes / sizeof(double), 0.0, h->stream);
    h->launches++;
    CUDA_CHECK_RET(cudaGetLastError());
    return 0;
}

extern "C" void *admpc_host_alloc(unsigned long long bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { admpc_set_error("cudaHostAlloc", "failed"); cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" int admpc_host_free(void *p) { CUDA_CHECK_RET(cudaFreeHost(p)); return 0; }

extern "C" int admpc_measure_fp64_peak(int device, double *tflops)
{
    if (!tflops) return ADMPC_E_ARG;
    const double v = run_fp64_peak(device, 0);
    if (v <= 0) { admpc_set_error("admpc_measure_fp64_peak", "CUDA failure"); return ADMPC_E_CUDA; }
    *tflops = v;
    return 0;
}
extern "C" int admpc_measure_fp64_mix(int device, int int_ops_per_8_dfma, double *tflops)
{
    if (!tflops || int_ops_per_8_dfma < 0 || (int_ops_per_8_dfma > 8 && int_ops_per_8_dfma != 16)) return ADMPC_E_ARG;
    const double v = run_fp64_peak(device, int_ops_per_8_dfma);
    if (v <= 0) { admpc_set_error("admpc_measure_fp64_mix", "CUDA failure"); return ADMPC_E_CUDA; }
    *tflops = v;
    return 0;
}

// ---------------------------------------------------------------------------------------------- reference generation ---
struct TrackHost { std::vector<double> dev; int L = 0, H = 0, stop = 0; };
int refgen_build_track(int L, const double *traj, int H, double dt, TrackHost &T);
void launch_refgen(const Params &P, const double *trk, int L, int H, double dt, int anchor, double *info, cudaStream_t s);

extern "C" int admpc_batch_set_track(admpc_batch *h, int L, const double *traj, int H, double traj_dt)
{
    if (h && h->P.o.model_variant != 0) { admpc_set_error("admpc_batch_set_track", "Cartesian model only (the Frenet variant takes its reference in path coordinates)"); return ADMPC_E_UNSUPPORTED; }
    if (!h) return ADMPC_E_ARG;
    if (H < h->P.o.N) { admpc_set_error("admpc_batch_set_track", "reference horizon H must be >= N (gp_ad_mpc_node.py:172-175 single-point branch is not supported)"); return ADMPC_E_UNSUPPORTED; }
    if (h->track_anchor && H > 64) { admpc_set_error("admpc_batch_set_track", "anchored mode supports H <= 64"); return ADMPC_E_UNSUPPORTED; }
    TrackHost T;
    int r = refgen_build_track(L, traj, H, traj_dt, T);
    if (r) { admpc_set_error("admpc_batch_set_track", "bad track (need L >= 2 rows of [vel,x,y,psi,cdist,curv], H >= 4)"); return r; }
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    const size_t bytes = T.dev.size() * sizeof(double);
    if (bytes > h->track_cap) {
        cudaFree(h->track);
        CUDA_CHECK_RET(cudaMalloc(&h->track, bytes));
        h->track_cap = bytes;
    }
    if (!h->track_info) CUDA_CHECK_RET(cudaMalloc(&h->track_info, (size_t)3 * h->P.Bp * sizeof(double)));
    CUDA_CHECK_RET(cudaMemcpyAsync(h->track, T.dev.data(), bytes, cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    h->track_L = L; h->track_H = H; h->track_stop = T.stop; h->track_dt = traj_dt;
    return 0;
}

extern "C" int admpc_batch_set_track_anchor(admpc_batch *h, int anchor_at_closest)
{
    if (!h) return ADMPC_E_ARG;
    if (anchor_at_closest && h->track_H > 64) { admpc_set_error("admpc_batch_set_track_anchor", "anchored mode supports H <= 64"); return ADMPC_E_UNSUPPORTED; }
    h->track_anchor = anchor_at_closest ? 1 : 0;
    return 0;
}

extern "C" int admpc_batch_make_yref(admpc_batch *h)
{
    if (!h) return ADMPC_E_ARG;
    if (!h->track) { admpc_set_error("admpc_batch_make_yref", "no track set"); return ADMPC_E_STATE; }
    if (h->track_anchor && h->track_H > 64) { admpc_set_error("admpc_batch_make_yref", "anchored mode supports H <= 64"); return ADMPC_E_UNSUPPORTED; }
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    launch_refgen(h->P, h->track, h->track_L, h->track_H, h->track_dt, h->track_anchor, h->track_info, h->stream);
    h->launches++;
    CUDA_CHECK_RET(cudaGetLastError());
    return 0;
}

// pose-only variant of solve_host: H2D(x0, p) -> make_yref on the device -> solve -> D2H(u, x, status); async.
extern "C" int admpc_batch_solve_pose_async(admpc_batch *h, const double *x0, const double *p_scalar, double *u_out,
                                            double *x_out, int *status_out)
{
    if (!h || !x0) return ADMPC_E_ARG;
    int r;
    if ((r = admpc_batch_set_x0(h, x0))) return r;
    if (p_scalar) { if ((r = admpc_batch_set_p_scalar(h, p_scalar))) return r; }
    else if (h->P.o.blend_max > h->P.o.blend_min) {        // no p from the host: blend from the measured v_x on the device
        launch_blend(h->P, h->stream);
        h->launches++;
    }
    if ((r = admpc_batch_make_yref(h))) return r;
    return solve_host_enqueue(h, nullptr, nullptr, nullptr, u_out, x_out, status_out);
}

extern "C" int admpc_batch_get_yref(admpc_batch *h, double *yref) { return get_rows(h, h->P.yref, yref, h->P.o.N * 9 + 7); }

extern "C" int admpc_batch_get_waypoint_info(admpc_batch *h, double *s0, double *e_y0, double *e_psi0, int *stop)
{
    if (!h || !h->track_info) return ADMPC_E_STATE;
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    const size_t n = (size_t)h->P.B * sizeof(double);
    if (s0) CUDA_CHECK_RET(cudaMemcpyAsync(s0, h->track_info, n, cudaMemcpyDeviceToHost, h->stream));
    if (e_y0) CUDA_CHECK_RET(cudaMemcpyAsync(e_y0, h->track_info + h->P.Bp, n, cudaMemcpyDeviceToHost, h->stream));
    if (e_psi0) CUDA_CHECK_RET(cudaMemcpyAsync(e_psi0, h->track_info + 2 * (size_t)h->P.Bp, n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    if (stop) *stop = h->track_stop;
    return 0;
}

// ---------------------------------------------------------------------------------------------- closed loop ------------
void launch_postsolve(const Params &P, double *prev_u, int *ibuf, double *u_apply, double *x_next, int threshold, int advance, cudaStream_t s);

static int loop_alloc(admpc_batch *h)
{
    if (h->loop_prev_u) return 0;
    const size_t Bp = h->P.Bp, N = h->P.o.N;
    CUDA_CHECK_RET(cudaMalloc(&h->loop_prev_u, (2 * N + 2 + 7) * Bp * sizeof(double)));
    CUDA_CHECK_RET(cudaMalloc(&h->loop_i, 4 * Bp * sizeof(int)));
    CUDA_CHECK_RET(cudaMemsetAsync(h->loop_prev_u, 0, (2 * N + 2 + 7) * Bp * sizeof(double), h->stream));
    CUDA_CHECK_RET(cudaMemsetAsync(h->loop_i, 0, 4 * Bp * sizeof(int), h->stream));
    return 0;
}

// validity check + backup control + safety counter (+ plant step when advance != 0) for the last solve
static int postsolve_impl(admpc_batch *h, int advance, int safe_threshold)
{
    if (h && h->P.o.model_variant != 0) { admpc_set_error("admpc_batch_postsolve", "Cartesian model only (the Frenet variant takes its reference in path coordinates)"); return ADMPC_E_UNSUPPORTED; }
    if (!h) return ADMPC_E_ARG;
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    int r = loop_alloc(h);
    if (r) return r;
    const size_t Bp = h->P.Bp, N = h->P.o.N;
    launch_postsolve(h->P, h->loop_prev_u, h->loop_i, h->loop_prev_u + 2 * N * Bp, h->loop_prev_u + (2 * N + 2) * Bp, safe_threshold, advance, h->stream);
    h->launches++;
    CUDA_CHECK_RET(cudaGetLastError());
    return 0;
}
extern "C" int admpc_batch_postsolve(admpc_batch *h, int advance, int safe_threshold)
{
    const int r = postsolve_impl(h, advance, safe_threshold);
    if (h) h->gat_fresh = false;      // the iterate moved after the last fused-gather epilogue
    return r;
}

// `steps` closed-loop control steps entirely on the device: [make_yref] -> solve -> postsolve(advance) ...
// log_x: optional host buffer [steps+1][B][7] receiving the plant state before every step and after the last one.
static int closed_loop_impl(admpc_batch *h, int steps, int use_track, int safe_threshold, double *log_x)
{
    if (h && h->P.o.model_variant != 0) { admpc_set_error("admpc_batch_closed_loop", "Cartesian model only (the Frenet variant takes its reference in path coordinates)"); return ADMPC_E_UNSUPPORTED; }
    if (!h || steps < 1) return ADMPC_E_ARG;
    if (use_track && !h->track) { admpc_set_error("admpc_batch_closed_loop", "no track set"); return ADMPC_E_STATE; }
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    int r = loop_alloc(h);
    if (r) return r;
    const size_t Bp = h->P.Bp, B = h->P.B;
    double *dlog = nullptr;
    if (log_x) CUDA_CHECK_RET(cudaMalloc(&dlog, (size_t)(steps + 1) * B * 7 * sizeof(double)));
    for (int t = 0; t < steps; t++) {
        if (dlog) { launch_transpose_out(h->P.x0, dlog + (size_t)t * B * 7, (int)B, (int)Bp, 7, h->stream); h->launches++; }
        if (use_track && (r = admpc_batch_make_yref(h))) break;
        if ((r = admpc_batch_solve(h))) break;
        if ((r = admpc_batch_postsolve(h, 1, safe_threshold))) break;
    }
    if (!r && dlog) {
        launch_transpose_out(h->P.x0, dlog + (size_t)steps * B * 7, (int)B, (int)Bp, 7, h->stream);
        h->launches++;
        cudaError_t e = cudaMemcpyAsync(log_x, dlog, (size_t)(steps + 1) * B * 7 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
        if (e != cudaSuccess) { admpc_set_error("closed_loop log copy", cudaGetErrorString(e)); r = ADMPC_E_CUDA; }
    }
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (dlog) cudaFree(dlog);
    if (r) return r;
    if (e != cudaSuccess) { admpc_set_error("admpc_batch_closed_loop", cudaGetErrorString(e)); return ADMPC_E_CUDA; }
    return 0;
}
extern "C" int admpc_batch_closed_loop(admpc_batch *h, int steps, int use_track, int safe_threshold, double *log_x)
{
    const int r = closed_loop_impl(h, steps, use_track, safe_threshold, log_x);
    if (h) h->gat_fresh = false;      // the iterate moved after the last fused-gather epilogue
    return r;
}

extern "C" int admpc_batch_get_loop_info(admpc_batch *h, int *valid, int *safe_count, int *cmd_ok, double *u_apply, double *x0)
{
    if (!h || !h->loop_prev_u) return ADMPC_E_STATE;
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    const size_t Bp = h->P.Bp, N = h->P.o.N, n = (size_t)h->P.B * sizeof(int);
    if (safe_count) CUDA_CHECK_RET(cudaMemcpyAsync(safe_count, h->loop_i + Bp, n, cudaMemcpyDeviceToHost, h->stream));
    if (valid) CUDA_CHECK_RET(cudaMemcpyAsync(valid, h->loop_i + 2 * Bp, n, cudaMemcpyDeviceToHost, h->stream));
    if (cmd_ok) CUDA_CHECK_RET(cudaMemcpyAsync(cmd_ok, h->loop_i + 3 * Bp, n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    int r = 0;
    if (u_apply) r = get_rows(h, h->loop_prev_u + 2 * N * Bp, u_apply, 2);
    if (!r && x0) r = get_rows(h, h->P.x0, x0, 7);
    return r;
}

// ---------------------------------------------------------------------------------------------- multi-GPU -------
extern "C" int admpc_nccl_unique_id(void *id128)
{
    int r = nccl_load();
    if (r) return r;
    NCCL_CHECK_RET(g_nccl.GetUniqueId((nccl_uid *)id128));
    return 0;
}
extern "C" int admpc_batch_comm_init(admpc_batch *h, const void *id128, int rank, int nranks)
{
    if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return ADMPC_E_ARG;
    int r = nccl_load();
    if (r) return r;
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    nccl_uid id;
    memcpy(&id, id128, sizeof id);
    NCCL_CHECK_RET(g_nccl.CommInitRank(&h->comm, nranks, id, rank));
    h->rank = rank; h->nranks = nranks;
    if (!h->bar_buf) { CUDA_CHECK_RET(cudaMalloc(&h->bar_buf, 16)); CUDA_CHECK_RET(cudaMemset(h->bar_buf, 0, 16)); }
    // the gathered layout is rank * block: every rank must hold the same number of instances (max B == min B)
    int bb[2] = {h->P.B, -h->P.B};
    CUDA_CHECK_RET(cudaMemcpyAsync(h->bar_buf + 2, bb, sizeof bb, cudaMemcpyHostToDevice, h->stream));
    NCCL_CHECK_RET(g_nccl.AllReduce(h->bar_buf + 2, h->bar_buf + 2, 2, NCCL_INT32, NCCL_MAX, h->comm, h->stream));
    CUDA_CHECK_RET(cudaMemcpyAsync(bb, h->bar_buf + 2, sizeof bb, cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    h->b_uniform = (bb[0] == -bb[1]);
    return 0;
}
extern "C" int admpc_batch_bcast_gp(admpc_batch *h, int root, int nout, int M, int dz, const int *feat, const int *rows,
                                    const double *X, const double *alpha, const double *ell, const double *sigma_f,
                                    const double *y_mean, int stage0_trigger)
{
    if (!h || !h->comm) { admpc_set_error("admpc_batch_bcast_gp", "communicator not initialised"); return ADMPC_E_STATE; }
    if (root < 0 || root >= h->nranks) return ADMPC_E_ARG;
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    // A collective: whatever happens locally, every rank takes part in the header broadcast, the agreement all-reduce
    // and (only if all ranks are fine) the blob broadcast, so that a local failure can never leave the peers blocked.
    int hdr[5 + ADMPC_DZMAX + ADMPC_GPOUT_MAX] = {0};
    int rc = 0;
    if (h->rank == root) {
        rc = upload_gp(h, 1, nout, M, dz, feat, rows, X, alpha, ell, sigma_f, y_mean, nullptr, stage0_trigger);   // single model
        if (rc == 0) rc = gp_res_alloc(h);
        hdr[0] = (rc == 0);
        if (rc == 0) {
            hdr[1] = nout; hdr[2] = M; hdr[3] = dz; hdr[4] = stage0_trigger;
            for (int d = 0; d < dz; d++) hdr[5 + d] = feat[d];
            for (int j = 0; j < nout; j++) hdr[5 + ADMPC_DZMAX + j] = rows[j];
        }
    }
    int *dh = h->stage_status;
    CUDA_CHECK_RET(cudaMemcpyAsync(dh, hdr, sizeof hdr, cudaMemcpyHostToDevice, h->stream));
    NCCL_CHECK_RET(g_nccl.Broadcast(dh, dh, sizeof hdr / sizeof(int), NCCL_INT32, root, h->comm, h->stream));
    CUDA_CHECK_RET(cudaMemcpyAsync(hdr, dh, sizeof hdr, cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    Params &P = h->P;
    int bad = hdr[0] ? 0 : 1;
    if (h->rank != root && hdr[0]) {
        // the same checks upload_gp applies on the root, on the sizes that arrived
        nout = hdr[1]; M = hdr[2]; dz = hdr[3];
        bool okv = nout >= 0 && nout <= ADMPC_GPOUT_MAX && (nout == 0 || (M > 0 && dz > 0 && dz <= ADMPC_DZMAX));
        for (int d = 0; okv && nout > 0 && d < dz; d++) okv = hdr[5 + d] >= 2 && hdr[5 + d] <= 8;
        for (int j = 0; okv && j < nout; j++) okv = hdr[5 + ADMPC_DZMAX + j] >= 3 && hdr[5 + ADMPC_DZMAX + j] <= 5;
        const size_t bytes = okv && nout > 0 ? gp_blob_doubles(nout, M, dz) * sizeof(double) : 0;
        if (okv && bytes > 220 * 1024) okv = false;
        if (!okv) { admpc_set_error("admpc_batch_bcast_gp", "received GP header is invalid / exceeds the 220 KB staging budget"); bad = 1; rc = ADMPC_E_UNSUPPORTED; }
        else if (nout == 0) { P.o.gp_enabled = 0; }
        else {
            if (bytes > h->gp_blob_cap) {
                cudaFree(h->gp_blob);
                h->gp_blob = nullptr; h->gp_blob_cap = 0; P.gp.blob = nullptr; P.o.gp_enabled = 0;
                if (cudaMalloc(&h->gp_blob, bytes) != cudaSuccess) {
                    cudaGetLastError();
                    h->gp_blob = nullptr;
                    admpc_set_error("admpc_batch_bcast_gp", "cudaMalloc of the GP blob failed");
                    bad = 1; rc = ADMPC_E_CUDA;
                } else h->gp_blob_cap = bytes;
            }
            if (!bad) {
                const size_t stride = gp_stride(M, dz);
                P.gp.blob = h->gp_blob; P.gp.bytes = (int)bytes; P.gp.stride_out = (int)stride;
                P.gp.n_models = 1; P.gp.model_doubles = (int)(stride * nout); P.gp.centroids = nullptr;
                CUDA_CHECK_RET(cudaMemsetAsync(h->gp_sel, 0, (size_t)P.Bp * sizeof(int), h->stream));
                P.o.gp_enabled = 1; P.o.gp_nout = nout; P.o.gp_M = M; P.o.gp_dz = dz; P.o.gp_stage0_trigger = hdr[4];
                for (int d = 0; d < dz; d++) P.o.gp_feat[d] = hdr[5 + d];
                for (int j = 0; j < nout; j++) P.o.gp_row[j] = hdr[5 + ADMPC_DZMAX + j];
                if ((rc = gp_res_alloc(h))) bad = 1;
            }
        }
    }
    // everybody or nobody
    CUDA_CHECK_RET(cudaMemcpyAsync(h->bar_buf + 1, &bad, sizeof bad, cudaMemcpyHostToDevice, h->stream));
    NCCL_CHECK_RET(g_nccl.AllReduce(h->bar_buf + 1, h->bar_buf + 1, 1, NCCL_INT32, NCCL_SUM, h->comm, h->stream));
    int nbad = 0;
    CUDA_CHECK_RET(cudaMemcpyAsync(&nbad, h->bar_buf + 1, sizeof nbad, cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    if (nbad) {
        P.o.gp_enabled = 0;
        if (rc == 0) { admpc_set_error("admpc_batch_bcast_gp", "another rank could not take the GP model"); rc = ADMPC_E_STATE; }
        return rc;
    }
    if (P.o.gp_enabled && P.gp.bytes > 0)
        NCCL_CHECK_RET(g_nccl.Broadcast(h->gp_blob, h->gp_blob, (size_t)P.gp.bytes / sizeof(double), NCCL_FLOAT64, root, h->comm, h->stream));
    CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    return 0;
}

// host copy of the blocks the last gather left on the root (rank-major, instance-major inside)
extern "C" int admpc_batch_get_gathered(admpc_batch *h, double *u_all, double *x_all, int *status_all)
{
    if (!h || !h->gpack) { admpc_set_error("admpc_batch_get_gathered", "nothing gathered on this rank"); return ADMPC_E_STATE; }
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    const Params &P = h->P;
    const size_t nu = (size_t)P.B * P.o.N * 2, nx = (size_t)P.B * (P.o.N + 1) * 7;
    const size_t bytes = gather_block_bytes(P);
    for (int r = 0; r < h->nranks; r++) {
        const char *blk = h->gpack + ((size_t)h->gat_last * h->nranks + r) * bytes;
        if (u_all) CUDA_CHECK_RET(cudaMemcpyAsync(u_all + (size_t)r * nu, blk, nu * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (x_all) CUDA_CHECK_RET(cudaMemcpyAsync(x_all + (size_t)r * nx, blk + nu * sizeof(double), nx * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (status_all) CUDA_CHECK_RET(cudaMemcpyAsync(status_all + (size_t)r * P.B, blk + (nu + nx) * sizeof(double), (size_t)P.B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    }
    CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    return 0;
}

// this rank's slice of the half of the root's block that is written next
static void gather_point(admpc_batch *h)
{
    Params &P = h->P;
    const size_t bytes = gather_block_bytes(P);
    const size_t nu = (size_t)P.B * P.o.N * 2, nx = (size_t)P.B * (P.o.N + 1) * 7;
    char *slice = ((h->rank == h->gat_root) ? h->gpack : h->gpack_peer) + ((size_t)h->gat_par * h->nranks + h->rank) * bytes;
    P.gat_u = (double *)slice; P.gat_x = P.gat_u + nu; P.gat_st = (int *)(P.gat_x + nx);
}

// Collective.  Sets up the FUSED gather towards `root`: the root allocates the gathered block and exports it through
// CUDA IPC (the 64-byte handle travels by ncclBroadcast), every other rank maps it; from then on the epilogue of the QP
// warp kernels writes each instance's [u | x | status] straight into the rank's slice of the root's block (stores over
// NVLink, overlapped with the solve) and admpc_batch_gather shrinks to a 4-byte all-reduce = the stream-ordered
// completion barrier.  Returns 1 when the fused path is active on all ranks, 0 when the NCCL send/recv path stays
// (IPC unavailable, ADMPC_GATHER=nccl), < 0 on error.
extern "C" int admpc_batch_gather_enable(admpc_batch *h, int root)
{
    if (!h || !h->comm) { admpc_set_error("admpc_batch_gather_enable", "communicator not initialised"); return ADMPC_E_STATE; }
    if (root < 0 || root >= h->nranks) return ADMPC_E_ARG;
    if (!h->b_uniform) { admpc_set_error("admpc_batch_gather_enable", "ranks hold different numbers of instances (pad the batch to a multiple of the rank count)"); return ADMPC_E_UNSUPPORTED; }
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    Params &P = h->P;
    const size_t bytes = gather_block_bytes(P);
    const char *env = getenv("ADMPC_GATHER");
    struct { cudaIpcMemHandle_t hd; int ok; int pad[3]; } msg;
    memset(&msg, 0, sizeof msg);
    msg.ok = !(env && !strcmp(env, "nccl"));
    if (h->rank == root) {
        if (!h->gpack) CUDA_CHECK_RET(cudaMalloc(&h->gpack, 2 * bytes * h->nranks));       // two halves, see gat_par
        if (msg.ok && cudaIpcGetMemHandle(&msg.hd, h->gpack) != cudaSuccess) { cudaGetLastError(); msg.ok = 0; }
    }
    char *d = nullptr;
    CUDA_CHECK_RET(cudaMalloc(&d, sizeof msg));
    CUDA_CHECK_RET(cudaMemcpyAsync(d, &msg, sizeof msg, cudaMemcpyHostToDevice, h->stream));
    NCCL_CHECK_RET(g_nccl.Broadcast(d, d, sizeof msg, NCCL_INT8, root, h->comm, h->stream));
    CUDA_CHECK_RET(cudaMemcpyAsync(&msg, d, sizeof msg, cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    int mine = msg.ok;
    if (mine && h->rank != root && !h->gpack_peer) {
        void *pp = nullptr;
        if (cudaIpcOpenMemHandle(&pp, msg.hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); mine = 0; }
        else h->gpack_peer = (char *)pp;
    }
    // everybody or nobody
    CUDA_CHECK_RET(cudaMemcpyAsync(d, &mine, sizeof mine, cudaMemcpyHostToDevice, h->stream));
    NCCL_CHECK_RET(g_nccl.AllReduce(d, d, 1, NCCL_INT32, NCCL_SUM, h->comm, h->stream));
    int total = 0;
    CUDA_CHECK_RET(cudaMemcpyAsync(&total, d, sizeof total, cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    cudaFree(d);
    if (total != h->nranks) {
        if (h->gpack_peer) { cudaIpcCloseMemHandle(h->gpack_peer); h->gpack_peer = nullptr; }
        h->gat_on = false; h->gat_fresh = false;
        P.gat_u = nullptr; P.gat_x = nullptr; P.gat_st = nullptr;
        return 0;
    }
    h->gat_on = true; h->gat_fresh = false; h->gat_root = root; h->gat_par = 0; h->gat_last = 0;
    gather_point(h);
    return 1;
}

extern "C" int admpc_batch_gather(admpc_batch *h, int root, double *u_all, double *x_all, int *status_all)
{
    if (!h || !h->comm) { admpc_set_error("admpc_batch_gather", "communicator not initialised"); return ADMPC_E_STATE; }
    if (!h->b_uniform) { admpc_set_error("admpc_batch_gather", "ranks hold different numbers of instances (pad the batch to a multiple of the rank count)"); return ADMPC_E_UNSUPPORTED; }
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    const Params &P = h->P;
    const int N = P.o.N, B = P.B;
    int half = 0;                                  // which half of the root's block this call delivers
    const size_t nu = (size_t)B * N * 2, nx = (size_t)B * (N + 1) * 7;
    const size_t bytes = gather_block_bytes(P);                                      // one packed block per rank
    const bool fused = h->gat_on && root == h->gat_root;
    if (fused) {
        // the QP kernel has already written this rank's slice of the root's block; if the iterate changed since (or a
        // kernel variant without the fused epilogue ran), pack it now -- same destination, plain stores over NVLink
        if (!h->gat_fresh) {
            launch_transpose_out(P.ub, P.gat_u, B, P.Bp, N * 2, h->stream);
            launch_transpose_out(P.xb, P.gat_x, B, P.Bp, (N + 1) * 7, h->stream);
            h->launches += 2;
            CUDA_CHECK_RET(cudaGetLastError());
            CUDA_CHECK_RET(cudaMemcpyAsync(P.gat_st, P.status, (size_t)B * sizeof(int), cudaMemcpyDefault, h->stream));
        }
        // stream-ordered completion: every rank enqueues the all-reduce after its stores, so once it has run on the
        // root's stream all slices have landed
        NCCL_CHECK_RET(g_nccl.AllReduce(h->bar_buf, h->bar_buf, 1, NCCL_INT32, NCCL_SUM, h->comm, h->stream));
        // later writes of this rank go to the other half (the half just delivered may still be read out by the root)
        half = h->gat_par;
        h->gat_par ^= 1;
        h->gat_fresh = false;
        gather_point(h);
    } else {
        if (!h->pack) CUDA_CHECK_RET(cudaMalloc(&h->pack, bytes));
        if (h->rank == root && !h->gpack) CUDA_CHECK_RET(cudaMalloc(&h->gpack, 2 * bytes * h->nranks));
        // instance-major [u | x | status] block, then ONE send per rank and one receive per peer on the root
        double *pu = (double *)h->pack, *px = pu + nu;
        int *pst = (int *)(px + nx);
        launch_transpose_out(P.ub, pu, B, P.Bp, N * 2, h->stream);
        launch_transpose_out(P.xb, px, B, P.Bp, (N + 1) * 7, h->stream);
        h->launches += 2;
        CUDA_CHECK_RET(cudaGetLastError());
        CUDA_CHECK_RET(cudaMemcpyAsync(pst, P.status, (size_t)B * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
        NCCL_CHECK_RET(g_nccl.GroupStart());
        if (h->rank != root) NCCL_CHECK_RET(g_nccl.Send(h->pack, bytes, NCCL_INT8, root, h->comm, h->stream));
        else {
            for (int r = 0; r < h->nranks; r++)
                if (r != root) NCCL_CHECK_RET(g_nccl.Recv(h->gpack + (size_t)r * bytes, bytes, NCCL_INT8, r, h->comm, h->stream));
        }
        NCCL_CHECK_RET(g_nccl.GroupEnd());
        if (h->rank == root) CUDA_CHECK_RET(cudaMemcpyAsync(h->gpack + (size_t)root * bytes, h->pack, bytes, cudaMemcpyDeviceToDevice, h->stream));
    }
    h->gat_last = half;
    if (h->rank == root && (u_all || x_all || status_all)) {
        for (int r = 0; r < h->nranks; r++) {
            const char *blk = h->gpack + ((size_t)half * h->nranks + r) * bytes;
            if (u_all) CUDA_CHECK_RET(cudaMemcpyAsync(u_all + (size_t)r * nu, blk, nu * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            if (x_all) CUDA_CHECK_RET(cudaMemcpyAsync(x_all + (size_t)r * nx, blk + nu * sizeof(double), nx * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            if (status_all) CUDA_CHECK_RET(cudaMemcpyAsync(status_all + (size_t)r * B, blk + (nu + nx) * sizeof(double), (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        }
        CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    }
    return 0;       // stream-ordered: admpc_batch_wait (or the next synchronising call) completes it
}

extern "C" int admpc_batch_barrier(admpc_batch *h)
{
    if (!h || !h->comm) return ADMPC_E_STATE;
    CUDA_CHECK_RET(cudaSetDevice(h->device));
    int *d = h->stage_status;
    NCCL_CHECK_RET(g_nccl.AllReduce(d, d, 1, NCCL_INT32, NCCL_SUM, h->comm, h->stream));
    CUDA_CHECK_RET(cudaStreamSynchronize(h->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------- acados shim -----
struct sim_car_solver_capsule {
    admpc_opts opts;
    bool opts_set = false;
    admpc_batch *h = nullptr;
    int N = 0;
    std::vector<double> x0, yref, p, kappa, x, u, pi, lam, t, sl, su;
    bool iterate_dirty = false, duals_stale = false, duals_dirty = false;
    double *hio = nullptr, *dio = nullptr;      // pinned host / device staging blocks of the single-instance fast path
    double *hio_dev = nullptr;                  // device-side address of the pinned block (zero-copy)
    bool zero_copy = false;
    // the fast path as a CUDA graph (one launch instead of ~12 stream calls): [0] without / [1] with a new initial guess; a graph
    // is replayed only while the kernel parameters it captured are byte-identical to the handle's current ones
    cudaGraphExec_t gexec[2] = {nullptr, nullptr};
    Params gparams[2];
    bool gprof[2] = {false, false};
    long long glaunches[2] = {0, 0};
    int rti_solves = 0;
    bool graph_ok = true;
    int status = 0, qp_status = 0, qp_iter = 0, sqp_iter = 1;
    bool nlp_sqp = false;            // nlp_solver_type: false "SQP_RTI" (shipped), true "SQP" (point-reference mode)
    int nlp_max_iter = 100;          // sim_car_acados_ocp.json:868
    double nlp_tol[4] = {1e-6, 1e-6, 1e-6, 1e-6};   // :870-873
    double time_tot = 0.0;
    double res[4] = {0, 0, 0, 0};
};

extern "C" sim_car_solver_capsule *sim_car_acados_create_capsule(void)
{
    sim_car_solver_capsule *c = new sim_car_solver_capsule();
    admpc_default_opts(&c->opts);
    c->opts.N = 40;     // SIM_CAR_N, acados_solver_sim_car.h:66
    return c;
}
extern "C" int sim_car_acados_free_capsule(sim_car_solver_capsule *c) { delete c; return 0; }
extern "C" int sim_car_acados_set_opts(sim_car_solver_capsule *c, const admpc_opts *o)
{
    if (!c || !o) return ADMPC_E_ARG;
    if (c->h) { admpc_set_error("sim_car_acados_set_opts", "solver already created"); return ADMPC_E_STATE; }
    c->opts = *o;
    c->opts_set = true;
    return 0;
}
// nlp_solver_type of the capsule: "SQP_RTI" (what the reference ships) or "SQP" (its point-reference mode,
// create_ros_ad_mpc.py:47-51); max_iter <= 0 / tol4 NULL keep the acados defaults (100, 1e-6).
extern "C" int sim_car_acados_set_nlp_solver(sim_car_solver_capsule *c, const char *type, int max_iter, const double *tol4)
{
    if (!c || !type) return ADMPC_E_ARG;
    if (!strcmp(type, "SQP")) c->nlp_sqp = true;
    else if (!strcmp(type, "SQP_RTI")) c->nlp_sqp = false;
    else { admpc_set_error("sim_car_acados_set_nlp_solver", "nlp_solver_type must be SQP_RTI or SQP"); return ADMPC_E_ARG; }
    if (max_iter > 0) c->nlp_max_iter = max_iter;
    if (tol4) memcpy(c->nlp_tol, tol4, sizeof c->nlp_tol);
    return 0;
}
extern "C" int sim_car_acados_create_with_discretization(sim_car_solver_capsule *c, int N, double *ts)
{
    if (!c) return ADMPC_E_ARG;
    if (c->h) { admpc_set_error("sim_car_acados_create", "already created"); return ADMPC_E_STATE; }
    admpc_opts o = c->opts;
    o.N = N;
    if (ts) {
        for (int k = 1; k < N; k++)
            if (fabs(ts[k] - ts[0]) > 1e-12 * fabs(ts[0])) { admpc_set_error("create_with_discretization", "non-uniform time steps are not supported"); return ADMPC_E_UNSUPPORTED; }
        o.dt = ts[0];
    }
    int r = admpc_batch_create(&o, 1, 0, &c->h);
    if (r) return r;
    c->opts = o;
    c->N = N;
    c->x0.assign(7, 0.0); c->yref.assign((size_t)N * 9 + 7, 0.0); c->p.assign(N, 0.0); c->kappa.assign(N, 0.0);
    c->x.assign((size_t)(N + 1) * 7, 0.0); c->u.assign((size_t)N * 2, 0.0); c->pi.assign((size_t)N * 7, 0.0);
    const size_t nc = (size_t)con_rows(o);          // 10 rows per stage (shipped set) or 12 (the Frenet variant's own, con_set = 1)
    c->lam.assign((size_t)N * nc, 0.0); c->t.assign((size_t)N * nc, 0.0); c->sl.assign((size_t)N * 2, 0.0); c->su.assign((size_t)N * 2, 0.0);
    return 0;
}
extern "C" int sim_car_acados_create(sim_car_solver_capsule *c) { return c ? sim_car_acados_create_with_discretization(c, c->opts.N, nullptr) : ADMPC_E_ARG; }
extern "C" int sim_car_acados_update_time_steps(sim_car_solver_capsule *c, int N, double *ts)
{
    if (!c || !c->h || !ts) return ADMPC_E_ARG;
    if (N != c->N) { admpc_set_error("update_time_steps", "N must equal the created horizon"); return ADMPC_E_ARG; }   // .h:131-133
    for (int k = 1; k < N; k++)
        if (fabs(ts[k] - ts[0]) > 1e-12 * fabs(ts[0])) return ADMPC_E_UNSUPPORTED;
    c->h->P.o.dt = ts[0];
    c->opts.dt = ts[0];
    return 0;
}
// acados_solver_sim_car.h:137.  The reference's generated body prints "no partial condensing solver is used" and exits
// (.c:810-816: the shipped solver is FULL_CONDENSING_HPIPM); this library never condenses (OCP-structured Riccati IPM),
// so the call is accepted and ignored.
extern "C" int sim_car_acados_update_qp_solver_cond_N(sim_car_solver_capsule *c, int qp_solver_cond_N)
{
    (void)qp_solver_cond_N;
    return c ? 0 : ADMPC_E_ARG;
}
extern "C" int sim_car_acados_update_params(sim_car_solver_capsule *c, int stage, double *value, int np)
{
    if (!c || !c->h || !value) return ADMPC_E_ARG;
    if (np != ADMPC_NP) { admpc_set_error("sim_car_acados_update_params", "wrong number of parameters (expected 1)"); return ADMPC_E_ARG; }  // .c:862-866
    if (stage < 0 || stage > c->N) return ADMPC_E_ARG;
    if (stage < c->N) c->p[stage] = value[0];     // the parameter of the terminal node does not enter the problem
    return 0;
}
extern "C" int sim_car_acados_reset(sim_car_solver_capsule *c, int)
{
    if (!c || !c->h) return ADMPC_E_STATE;
    std::fill(c->x.begin(), c->x.end(), 0.0); std::fill(c->u.begin(), c->u.end(), 0.0);
    std::fill(c->pi.begin(), c->pi.end(), 0.0); std::fill(c->lam.begin(), c->lam.end(), 0.0);
    std::fill(c->t.begin(), c->t.end(), 0.0); std::fill(c->sl.begin(), c->sl.end(), 0.0); std::fill(c->su.begin(), c->su.end(), 0.0);
    c->iterate_dirty = false;
    return admpc_batch_reset(c->h);
}
extern "C" int sim_car_acados_free(sim_car_solver_capsule *c)
{
    if (!c) return ADMPC_E_ARG;
    for (int g = 0; g < 2; g++) if (c->gexec[g]) { cudaGraphExecDestroy(c->gexec[g]); c->gexec[g] = nullptr; }
    if (c->hio) { cudaFreeHost(c->hio); c->hio = nullptr; }
    if (c->dio) { cudaFree(c->dio); c->dio = nullptr; }
    int r = c->h ? admpc_batch_free(c->h) : 0;
    c->h = nullptr;
    return r;
}

extern "C" int sim_car_acados_set(sim_car_solver_capsule *c, int stage, const char *field, const double *v, int n)
{
    if (!c || !c->h || !field || !v) return ADMPC_E_ARG;
    const int N = c->N;
    if (stage < 0 || stage > N) { admpc_set_error("sim_car_acados_set", "stage out of range"); return ADMPC_E_ARG; }
    if (!strcmp(field, "yref")) {
        const int need = stage < N ? 9 : 7;
        if (n != need) { admpc_set_error("sim_car_acados_set", "yref has wrong length"); return ADMPC_E_ARG; }
        memcpy(&c->yref[(size_t)stage * 9], v, sizeof(double) * need);
        return 0;
    }
    if (!strcmp(field, "lbx") || !strcmp(field, "ubx")) {
        if (stage == 0 && n == 7) { memcpy(c->x0.data(), v, sizeof(double) * 7); return 0; }   // x0 via lbx=ubx (ad_3d_optimizer.py:441-442)
        admpc_set_error("sim_car_acados_set", "only the stage-0 state bound (x0, 7 values) can be changed per solve");
        return ADMPC_E_UNSUPPORTED;
    }
    if (!strcmp(field, "p")) { return sim_car_acados_update_params(c, stage, (double *)v, n); }
    if (!strcmp(field, "kappa")) {       // Frenet variant: path curvature at this shooting node
        if (n != 1 || stage >= N) return ADMPC_E_ARG;
        if (c->opts.model_variant != 1) { admpc_set_error("sim_car_acados_set", "kappa needs model_variant = 1 (Frenet)"); return ADMPC_E_STATE; }
        c->kappa[stage] = v[0]; return 0;
    }
    if (!strcmp(field, "x")) {
        if (n != 7) return ADMPC_E_ARG;
        memcpy(&c->x[(size_t)stage * 7], v, sizeof(double) * 7); c->iterate_dirty = true; return 0;
    }
    if (!strcmp(field, "u")) {
        if (n != 2 || stage >= N) return ADMPC_E_ARG;
        memcpy(&c->u[(size_t)stage * 2], v, sizeof(double) * 2); c->iterate_dirty = true; return 0;
    }
    // multipliers / slacks of the iterate (acados load_iterate, ad_3d_optimizer.py:454): same layouts as sim_car_acados_get
    if (!strcmp(field, "pi") || !strcmp(field, "sl") || !strcmp(field, "su") || !strcmp(field, "lam") || !strcmp(field, "t")) {
        if (stage >= N) return n == 0 ? 0 : ADMPC_E_ARG;               // nothing lives at the terminal node
        if (c->duals_stale) {                                           // start from what the device holds
            int r;
            if ((r = admpc_batch_get_pi(c->h, c->pi.data()))) return r;
            if ((r = admpc_batch_get_lam(c->h, c->lam.data()))) return r;
            if ((r = admpc_batch_get_t(c->h, c->t.data()))) return r;
            if ((r = admpc_batch_get_slacks(c->h, c->sl.data(), c->su.data()))) return r;
            c->duals_stale = false;
        }
        const int nc = con_rows(c->opts);
        const bool own = (nc == 12);                // con_set = 1: rows [lb u0 u1 e_y delta | ub .. | ls0 ls1 | us0 us1], slack 1 belongs to delta
        if (!strcmp(field, "pi")) { if (n != 7) return ADMPC_E_ARG; memcpy(&c->pi[(size_t)stage * 7], v, 56); }
        else if (!strcmp(field, "sl") || !strcmp(field, "su")) {
            double *d = ((field[1] == 'l') ? c->sl : c->su).data() + (size_t)stage * 2;
            if (own && stage == 0) { if (n != 1) return ADMPC_E_ARG; d[0] = v[0]; d[1] = 0.0; }      // no state slack at stage 0
            else { if (n != 2) return ADMPC_E_ARG; memcpy(d, v, 16); }
        } else {
            double *d = ((field[0] == 'l') ? c->lam : c->t).data() + (size_t)stage * nc;
            const double off = (field[0] == 'l') ? 0.0 : 1.0;             // rows that do not exist at stage 0 (x0 is eliminated)
            if (stage >= 1) { if (n != nc) return ADMPC_E_ARG; memcpy(d, v, sizeof(double) * nc); }
            else if (!own) {    // stage 0 arrives in the acados layout [lbu(2) lbx0(7) | ubu(2) ubx0(7) | ls(2) | us(2)]
                if (n != 22) return ADMPC_E_ARG;
                d[0] = v[0]; d[1] = v[1]; d[3] = v[9]; d[4] = v[10];
                for (int j = 0; j < 4; j++) d[6 + j] = v[18 + j];
                d[2] = off; d[5] = off;
            } else {            // con_set = 1, ad_mpc/debug.json lam_0: [lbu(2) lbx0(7) | ubu(2) ubx0(7) | ls(1) | us(1)]
                if (n != 20) return ADMPC_E_ARG;
                d[0] = v[0]; d[1] = v[1]; d[4] = v[9]; d[5] = v[10]; d[8] = v[18]; d[10] = v[19];
                d[2] = d[3] = d[6] = d[7] = d[9] = d[11] = off;
            }
        }
        c->duals_dirty = true;
        return 0;
    }
    admpc_set_error("sim_car_acados_set", "unknown field");
    return ADMPC_E_ARG;
}

extern "C" int sim_car_acados_solve(sim_car_solver_capsule *c)
{
    if (!c || !c->h) { admpc_set_error("sim_car_acados_solve", "solver not created"); return ADMPC_E_STATE; }
    admpc_batch *h = c->h;
    int r;
    if (c->duals_dirty) {       // multipliers set by the caller (load_iterate): ship them before the solve
        if ((r = admpc_batch_set_duals(h, c->pi.data(), c->lam.data(), c->t.data(), c->sl.data(), c->su.data()))) return r;
        c->duals_dirty = false;
    }
    if (!c->nlp_sqp) {
        // RTI fast path: ONE packed host->device block, ONE packed device->host block, one synchronisation
        const int N = c->N, nyr = 9 * N + 7, nx = (N + 1) * 7, nu = 2 * N;
        const size_t nin = (size_t)7 + nyr + N + N + nx + nu, nout = (size_t)nx + nu + 7;
        if (!c->hio) {
            CUDA_CHECK_RET(cudaSetDevice(h->device));
            CUDA_CHECK_RET(cudaHostAlloc((void **)&c->hio, (nin + nout) * sizeof(double), cudaHostAllocMapped));
            CUDA_CHECK_RET(cudaMalloc(&c->dio, (nin + nout) * sizeof(double)));
            c->zero_copy = !getenv("ADMPC_NO_ZEROCOPY") && cudaHostGetDevicePointer((void **)&c->hio_dev, c->hio, 0) == cudaSuccess;
            if (!c->zero_copy) cudaGetLastError();
        }
        double *hi = c->hio, *ho = c->hio + nin;
        memcpy(hi, c->x0.data(), 7 * sizeof(double));
        memcpy(hi + 7, c->yref.data(), nyr * sizeof(double));
        memcpy(hi + 7 + nyr, c->p.data(), N * sizeof(double));
        memcpy(hi + 7 + nyr + N, c->kappa.data(), N * sizeof(double));
        const bool wit = c->iterate_dirty;
        if (wit) {
            memcpy(hi + 7 + nyr + 2 * N, c->x.data(), nx * sizeof(double));
            memcpy(hi + 7 + nyr + 2 * N + nx, c->u.data(), nu * sizeof(double));
        }
        CUDA_CHECK_RET(cudaSetDevice(h->device));
        if ((r = admpc_batch_timer_start(h))) return r;
        // zero-copy: the pinned block is mapped into the device's address space (unified addressing), so the scatter kernel reads
        // its 2 KB straight from host memory and the gather kernel writes its 1.5 KB straight back -- two copy operations (and
        // their ~6 us each of launch latency) less on a path whose whole budget is 0.1 ms
        const bool zc = c->zero_copy;
        auto enqueue = [&]() -> int {
            if (!zc) CUDA_CHECK_RET(cudaMemcpyAsync(c->dio, hi, (wit ? nin : (size_t)7 + nyr + 2 * N) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            launch_capsule_scatter(h->P, zc ? c->hio_dev : c->dio, c->opts.model_variant == 1, wit, h->stream);
            if (int rr = admpc_batch_solve(h)) return rr;
            launch_capsule_gather(h->P, zc ? c->hio_dev + nin : c->dio + nin, h->stream);
            h->launches += 2;
            if (!zc) CUDA_CHECK_RET(cudaMemcpyAsync(ho, c->dio + nin, nout * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            return 0;
        };
        c->iterate_dirty = false;
        const int gi = wit ? 1 : 0;
        if (c->rti_solves == 0 && getenv("ADMPC_NO_GRAPH")) c->graph_ok = false;
        if (c->gexec[gi] && (memcmp(&c->gparams[gi], &h->P, sizeof(Params)) != 0 || c->gprof[gi] != h->profiling)) {
            cudaGraphExecDestroy(c->gexec[gi]);                 // something the kernels see has changed since the capture
            c->gexec[gi] = nullptr;
        }
        if (c->gexec[gi]) {
            CUDA_CHECK_RET(cudaGraphLaunch(c->gexec[gi], h->stream));
            h->launches += c->glaunches[gi];
        } else if (c->graph_ok && c->rti_solves >= 2) {         // (the first solves run eagerly: one-time kernel attributes, warm-up)
            const long long l0 = h->launches;
            cudaGraph_t graph = nullptr;
            bool captured = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
            if (captured) {
                r = enqueue();
                captured = cudaStreamEndCapture(h->stream, &graph) == cudaSuccess && r == 0 && graph != nullptr;
            }
            if (captured && cudaGraphInstantiate(&c->gexec[gi], graph, 0) == cudaSuccess) {
                c->gparams[gi] = h->P; c->gprof[gi] = h->profiling; c->glaunches[gi] = h->launches - l0;
                cudaGraphDestroy(graph);
                CUDA_CHECK_RET(cudaGraphLaunch(c->gexec[gi], h->stream));
            } else {                                            // capture not possible here: stay on the eager path for good
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                c->gexec[gi] = nullptr; c->graph_ok = false;
                h->launches = l0;
                if ((r = enqueue())) return r;
            }
        } else {
            if ((r = enqueue())) return r;
        }
        c->rti_solves++;
        float ms = 0;
        if ((r = admpc_batch_timer_stop(h, &ms))) return r;          // synchronises the stream
        c->time_tot = ms * 1e-3;
        memcpy(c->x.data(), ho, nx * sizeof(double));
        memcpy(c->u.data(), ho + nx, nu * sizeof(double));
        memcpy(c->res, ho + nx + nu, 4 * sizeof(double));
        c->status = (int)ho[nx + nu + 4]; c->qp_status = (int)ho[nx + nu + 5]; c->qp_iter = (int)ho[nx + nu + 6];
        c->duals_stale = true;
        c->sqp_iter = 1;
        return c->status;
    }
    if ((r = admpc_batch_timer_start(h))) return r;
    if ((r = admpc_batch_set_x0(h, c->x0.data()))) return r;
    if ((r = admpc_batch_set_yref(h, c->yref.data()))) return r;
    if ((r = admpc_batch_set_p(h, c->p.data()))) return r;
    if (c->opts.model_variant == 1 && (r = admpc_batch_set_kappa(h, c->kappa.data()))) return r;
    if (c->iterate_dirty) { if ((r = admpc_batch_set_iterate(h, c->x.data(), c->u.data()))) return r; c->iterate_dirty = false; }
    if (c->nlp_sqp) {
        if ((r = admpc_batch_solve_sqp(h, c->nlp_max_iter, c->nlp_tol, nullptr))) return r;     // Cartesian model only
    } else {
        if ((r = admpc_batch_solve(h))) return r;
    }
    float ms = 0;
    if ((r = admpc_batch_timer_stop(h, &ms))) return r;
    c->time_tot = ms * 1e-3;
    // primal solution and statuses now; multipliers / slacks are fetched on first use (most callers only read x, u)
    if ((r = admpc_batch_get_x(h, c->x.data()))) return r;
    if ((r = admpc_batch_get_u(h, c->u.data()))) return r;
    if ((r = admpc_batch_get_status(h, &c->status, &c->qp_status, &c->qp_iter))) return r;
    c->duals_stale = true;
    c->sqp_iter = 1;
    if (c->nlp_sqp) {
        if ((r = admpc_batch_get_sqp_info(h, &c->status, &c->sqp_iter, c->res))) return r;
    } else {
        CUDA_CHECK_RET(cudaMemcpy2D(c->res, sizeof(double), h->P.res_out, (size_t)h->P.Bp * sizeof(double), sizeof(double), 4, cudaMemcpyDeviceToHost));
    }
    return c->status;
}

extern "C" int sim_car_acados_get(sim_car_solver_capsule *c, int stage, const char *field, double *out, int n)
{
    if (!c || !c->h || !field || !out) return ADMPC_E_ARG;
    const int N = c->N;
    if (stage < 0 || stage > N) return ADMPC_E_ARG;
    if (!strcmp(field, "x")) { if (n != 7) return ADMPC_E_ARG; memcpy(out, &c->x[(size_t)stage * 7], 56); return 0; }
    if (stage >= N) { admpc_set_error("sim_car_acados_get", "field has no entry at the terminal node"); return ADMPC_E_ARG; }
    if (!strcmp(field, "u")) { if (n != 2) return ADMPC_E_ARG; memcpy(out, &c->u[(size_t)stage * 2], 16); return 0; }
    if (c->duals_stale) {
        int r;
        if ((r = admpc_batch_get_pi(c->h, c->pi.data()))) return r;
        if ((r = admpc_batch_get_lam(c->h, c->lam.data()))) return r;
        if ((r = admpc_batch_get_t(c->h, c->t.data()))) return r;
        if ((r = admpc_batch_get_slacks(c->h, c->sl.data(), c->su.data()))) return r;
        c->duals_stale = false;
    }
    if (!strcmp(field, "pi")) { if (n != 7) return ADMPC_E_ARG; memcpy(out, &c->pi[(size_t)stage * 7], 56); return 0; }
    const int nc = con_rows(c->opts);
    const bool own = (nc == 12);
    if (!strcmp(field, "sl") || !strcmp(field, "su")) {
        const double *s = ((field[1] == 'l') ? c->sl : c->su).data() + (size_t)stage * 2;
        if (own && stage == 0) { if (n != 1) return ADMPC_E_ARG; out[0] = s[0]; return 0; }           // debug.json: sl_0 has one entry
        if (n != 2) return ADMPC_E_ARG;
        memcpy(out, s, 16);
        return 0;
    }
    if (!strcmp(field, "lam") || !strcmp(field, "t")) {
        const std::vector<double> &src = (field[0] == 'l') ? c->lam : c->t;
        const double *s = &src[(size_t)stage * nc];
        if (stage >= 1) { if (n != nc) return ADMPC_E_ARG; memcpy(out, s, sizeof(double) * nc); return 0; }
        if (!own) {
            // stage 0 in acados layout: [lbu(2) lbx0(7) | ubu(2) ubx0(7) | ls(2) | us(2)]  (sim_car_iterate.json lam_0)
            if (n != 22) return ADMPC_E_ARG;
            for (int j = 0; j < 22; j++) out[j] = 1e-16;
            out[0] = s[0]; out[1] = s[1]; out[9] = s[3]; out[10] = s[4];
            for (int j = 0; j < 4; j++) out[18 + j] = s[6 + j];
            return 0;
        }
        // con_set = 1: [lbu(2) lbx0(7) | ubu(2) ubx0(7) | ls(1) | us(1)]  (ad_mpc/debug.json lam_0, 20 entries)
        if (n != 20) return ADMPC_E_ARG;
        for (int j = 0; j < 20; j++) out[j] = 1e-16;
        out[0] = s[0]; out[1] = s[1]; out[9] = s[4]; out[10] = s[5]; out[18] = s[8]; out[19] = s[10];
        return 0;
    }
    admpc_set_error("sim_car_acados_get", "unknown field");
    return ADMPC_E_ARG;
}

extern "C" int sim_car_acados_get_stat(sim_car_solver_capsule *c, const char *name, void *out)
{
    if (!c || !name || !out) return ADMPC_E_ARG;
    if (!strcmp(name, "sqp_iter")) { *(int *)out = c->sqp_iter; return 0; }   // RTI: exactly one SQP iteration
    if (!strcmp(name, "qp_iter")) { *(int *)out = c->qp_iter; return 0; }
    if (!strcmp(name, "qp_stat")) { *(int *)out = c->qp_status; return 0; }
    if (!strcmp(name, "status")) { *(int *)out = c->status; return 0; }
    if (!strcmp(name, "time_tot")) { *(double *)out = c->time_tot; return 0; }
    if (!strcmp(name, "kkt_norm_inf")) {
        double m = 0;
        for (double r : c->res) m = r > m ? r : m;
        *(double *)out = m;
        return 0;
    }
    admpc_set_error("sim_car_acados_get_stat", "unknown statistic");
    return ADMPC_E_ARG;
}

extern "C" void sim_car_acados_print_stats(sim_car_solver_capsule *c)
{
    if (!c) return;
    // same columns as acados_solver_sim_car.c:950-977
    printf("\niter\tqp_stat\tqp_iter\n");
    printf("%d\t%d\t%d\n", 1, c->qp_status, c->qp_iter);
}
