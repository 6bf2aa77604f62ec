// C-ABI entry points (include/deepsir_b200.h).  Each function validates its arguments, carves the
// caller's workspace, and enqueues kernels on the caller's stream.  No host synchronisation, no
// allocation, no global state besides a thread-local "last CUDA error" string.
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "graph.cuh"
#include "kabsch.cuh"
#include "keypoint.cuh"
#include "metrics.cuh"
#include "knn.cuh"
#include "match.cuh"
#include "match_tc.cuh"

namespace dsir {
static thread_local cudaError_t g_last_err = cudaSuccess;
void set_last_cuda_error(cudaError_t e) { g_last_err = e; }
static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- in-situ profiler (diagnostic) ----
bool g_profile_on = false;
namespace {
struct ProfMark { cudaEvent_t ev; const char *file; int line; };
std::vector<ProfMark> g_marks;
std::mutex g_prof_mu;
}  // namespace
void profile_mark(cudaStream_t st, const char *file, int line) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (g_marks.size() >= 65536) return;
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, st);
    g_marks.push_back({ev, file, line});
}
}  // namespace dsir

using namespace dsir;

extern "C" {

int dsir_version(void) { return 100; }

const char *dsir_strerror(int code) {
    switch (code) {
        case DSIR_OK: return "ok";
        case DSIR_ERR_BAD_ARG: return "bad argument (null pointer, non-positive size or unknown enum)";
        case DSIR_ERR_UNSUPPORTED: return "shape not supported by the sm_100a kernels";
        case DSIR_ERR_WORKSPACE: return "workspace missing or too small";
        case DSIR_ERR_CUDA: return "CUDA call failed (see dsir_last_cuda_error)";
        case DSIR_ERR_KNN_TOO_FEW: return "knn: fewer support points than k";
        case DSIR_ERR_NO_DEVICE: return "no sm_100 (B200) device: deepsir_b200 has no other code path";
        default: return "unknown error";
    }
}

const char *dsir_last_cuda_error(void) { return cudaGetErrorString(g_last_err); }

uint64_t dsir_launch_count(void) { return (uint64_t)g_launches.load(std::memory_order_relaxed); }

int dsir_device_check(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return DSIR_ERR_NO_DEVICE; }
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
        cudaGetLastError();
        return DSIR_ERR_NO_DEVICE;
    }
    return major == 10 ? DSIR_OK : DSIR_ERR_NO_DEVICE;
}

int dsir_profile_begin(dsir_stream_t stream) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto &m : g_marks) cudaEventDestroy(m.ev);
    g_marks.clear();
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) != cudaSuccess) return DSIR_ERR_CUDA;
    cudaEventRecord(ev, (cudaStream_t)stream);
    g_marks.push_back({ev, "begin", 0});
    g_profile_on = true;
    return DSIR_OK;
}

int dsir_profile_report(char *buf, size_t buf_bytes) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_profile_on = false;
    if (!buf || buf_bytes == 0) return DSIR_ERR_BAD_ARG;
    cudaDeviceSynchronize();
    std::map<std::string, std::pair<int, double>> agg;
    std::vector<std::string> order;
    double total = 0.0;
    for (size_t i = 1; i < g_marks.size(); ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, g_marks[i - 1].ev, g_marks[i].ev) != cudaSuccess) { cudaGetLastError(); continue; }
        const char *f = strrchr(g_marks[i].file, '/');
        std::string key = std::string(f ? f + 1 : g_marks[i].file) + ":" + std::to_string(g_marks[i].line);
        if (!agg.count(key)) order.push_back(key);
        agg[key].first += 1;
        agg[key].second += ms * 1e3;
        total += ms * 1e3;
    }
    std::string out;
    char line[256];
    for (auto &k : order) {
        snprintf(line, sizeof line, "%-28s launches %5d  total %10.1f us  %5.1f%%\n", k.c_str(), agg[k].first, agg[k].second,
                 total > 0 ? 100.0 * agg[k].second / total : 0.0);
        out += line;
    }
    snprintf(line, sizeof line, "%-28s launches %5d  total %10.1f us\n", "TOTAL", (int)g_marks.size() - 1, total);
    out += line;
    for (auto &m : g_marks) cudaEventDestroy(m.ev);
    g_marks.clear();
    strncpy(buf, out.c_str(), buf_bytes - 1);
    buf[buf_bytes - 1] = 0;
    return DSIR_OK;
}

/* ------------------------------------------------------------------ KNN ------------------------ */
namespace {

// Fork/join helper for the independent query launches of one pyramid.  The levels of a pyramid only share the grids
// (built first, on the caller's stream); the big level-0 self-kNN stays on the caller's stream, the coarse levels and
// the 1-NN up-sampling queries go to library-owned auxiliary streams so that their small grids fill the tail of the big
// launch instead of running one after the other.  Per host thread and device: 4 non-blocking streams used round-robin
// and a ring of timing-less events.  Everything forked is joined back before the entry point returns, so the caller
// still sees ONE stream (also under CUDA-graph capture, where fork/join by events is the supported pattern).
struct AuxStreams {
    static constexpr int NS = 4, NE = 32;
    int dev = -1;
    cudaStream_t s[NS] = {};
    cudaEvent_t e[NE] = {};
    int next_s = 0, next_e = 0;
    bool ready(int device) {
        if (dev == device) return true;
        if (dev != -1) return false;   // one device per host thread (one process per GPU); otherwise stay single-stream
        for (int i = 0; i < NS; ++i)
            if (cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return false; }
        for (int i = 0; i < NE; ++i)
            if (cudaEventCreateWithFlags(&e[i], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return false; }
        dev = device;
        return true;
    }
    cudaStream_t stream() { cudaStream_t r = s[next_s]; next_s = (next_s + 1) % NS; return r; }
    cudaEvent_t event() { cudaEvent_t r = e[next_e]; next_e = (next_e + 1) % NE; return r; }
};
thread_local AuxStreams g_aux;
const bool KNN_FORK = getenv("DSIR_KNN_NOFORK") == nullptr;

struct GridSlot {
    KnnGridHeader *hdr;
    int *cell_start;
    float4 *sorted;
    int n;
};

size_t grid_slot_bytes(int B, int n) {
    return ws_block((size_t)B * sizeof(KnnGridHeader)) + ws_block((size_t)B * (KNN_GRID_GMAX + 1) * sizeof(int)) +
           ws_block((size_t)B * n * sizeof(float4));
}

bool take_grid_slot(Workspace &W, int B, int n, GridSlot *s) {
    s->hdr = W.take<KnnGridHeader>((size_t)B);
    s->cell_start = W.take<int>((size_t)B * (KNN_GRID_GMAX + 1));
    s->sorted = W.take<float4>((size_t)B * n);
    s->n = n;
    return W.ok();
}

// spatial structure of one support cloud: brute force (none), uniform grid (knn_grid.cu), bucket tree (knn_tree.cu)
enum { SP_BRUTE = 0, SP_GRID = 1, SP_TREE = 2 };
// `Nmax` = the largest cloud that has to be indexed in the same call (support and queries share one structure type)
int knn_structure(int algo, int Ns, int Nmax, int k, bool pyramid = false) {
    if (algo == DSIR_KNN_BRUTE) return SP_BRUTE;
    if (algo == DSIR_KNN_GRID) return SP_GRID;
    if (algo == DSIR_KNN_TREE) return k <= KNN_TREE_MAX_K ? SP_TREE : SP_GRID;   // the caller checks the size limit
    if (Ns < KNN_GRID_MIN_POINTS) return SP_BRUTE;
    // AUTO.  Measured on 32 C2 clouds (profiles/knn_tree_r2.md): one k = 16 self query costs 0.53 ms through the bucket tree
    // (0.13 ms of it the kd-ordered build) against 0.47 ms through the grid, but a whole pyramid - whose 1-NN up-sampling
    // queries and coarse levels re-use the trees - 0.67 ms against 0.72 ms, and the pair of pyramids of a step 1.11 against
    // 1.28 ms: pyramids take the tree, single queries the grid.
    if (pyramid && Nmax <= KNN_TREE_MAX_POINTS && k <= KNN_TREE_MAX_K) return SP_TREE;
    return SP_GRID;
}

// tunables (overridable for experiments through DSIR_GRID_CPP / DSIR_GRID_R0)
float env_float(const char *name, float dflt) {
    const char *v = getenv(name);
    return v ? (float)atof(v) : dflt;
}
const float GRID_CELLS_PER_POINT = env_float("DSIR_GRID_CPP", 0.7f);
const float GRID_R0_CELLS = env_float("DSIR_GRID_R0", 1.0f);

}  // namespace

size_t dsir_knn_workspace_bytes(int B, int Ns, int Nq, int k, int algo) {
    if (B <= 0 || Ns <= 0) return 256;
    size_t bytes = ws_block((size_t)B * Ns * sizeof(float4)) + 256;
    const int sp = knn_structure(algo, Ns, Ns > Nq ? Ns : Nq, k);
    if (sp == SP_GRID) bytes += grid_slot_bytes(B, Ns) + (Nq > 0 ? grid_slot_bytes(B, Nq) + ws_block((size_t)B * Nq * sizeof(float4)) : 0);
    if (sp == SP_TREE) bytes += knn_tree_slot_bytes(B, Ns) + (Nq > 0 ? knn_tree_slot_bytes(B, Nq) + ws_block((size_t)B * Nq * sizeof(float4)) : 0);
    return bytes;
}

int dsir_knn_xyz(const float *support, int sup_stride, const float *query, int qry_stride, int B, int Ns, int Nq,
                 int k, int64_t *idx, float *dist2, void *ws, size_t ws_bytes, int algo, dsir_stream_t stream) {
    if (!support || !query || !idx || B <= 0 || Ns <= 0 || Nq < 0 || sup_stride < 3 || qry_stride < 3 || k <= 0)
        return DSIR_ERR_BAD_ARG;
    if (algo < DSIR_KNN_AUTO || algo > DSIR_KNN_TREE) return DSIR_ERR_BAD_ARG;
    if (k > 32) return DSIR_ERR_UNSUPPORTED;
    if (Ns < k) return DSIR_ERR_KNN_TOO_FEW;
    const int sp = knn_structure(algo, Ns, Ns > Nq ? Ns : Nq, k);
    if (sp == SP_TREE && (Ns > KNN_TREE_MAX_POINTS || Nq > KNN_TREE_MAX_POINTS)) return DSIR_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    Workspace W(ws, ws_bytes);
    float4 *sup4 = W.take<float4>((size_t)B * Ns);
    if (!W.ok()) return DSIR_ERR_WORKSPACE;
    int rc = launch_pack_xyz4(support, sup_stride, (long long)B * Ns, sup4, st);
    if (rc) return rc;
    if (sp == SP_TREE) {
        // bucket tree over the support; the queries get their own tree so that every warp owns 32 neighbouring queries
        // (for a self-query the two trees coincide)
        const bool self = (query == support) && Nq == Ns && qry_stride == sup_stride;
        if (Nq == 0) return DSIR_OK;
        KnnTreeView ts, tq;
        if (!knn_tree_take_slot(W, B, Ns, &ts)) return DSIR_ERR_WORKSPACE;
        if ((rc = launch_knn_tree_build(sup4, Ns, &ts, 1, B, st))) return rc;
        tq = ts;
        if (!self) {
            float4 *q4 = W.take<float4>((size_t)B * Nq);
            if (!W.ok() || !knn_tree_take_slot(W, B, Nq, &tq)) return DSIR_ERR_WORKSPACE;
            if ((rc = launch_pack_xyz4(query, qry_stride, (long long)B * Nq, q4, st))) return rc;
            if ((rc = launch_knn_tree_build(q4, Nq, &tq, 1, B, st))) return rc;
        }
        KnnTreeQueryParams Q{};
        Q.sup = ts; Q.qry = tq; Q.self = self ? 1 : 0; Q.k = k;
        Q.idx = idx; Q.idx_bs = (long long)Nq * k; Q.dist2 = dist2;
        return launch_knn_tree_query(Q, B, st);
    }
    if (sp == SP_BRUTE) {
        KnnBruteParams P{};
        P.sup4 = sup4; P.sup_bs = Ns;
        P.query = query; P.qry_bs = (long long)Nq * qry_stride; P.qry_stride = qry_stride;
        P.Ns = Ns; P.Nq = Nq; P.k = k;
        P.idx = idx; P.idx_bs = (long long)Nq * k; P.dist2 = dist2;
        return launch_knn_brute(P, B, st);
    }
    // grid over the support; the queries are binned into their own grid as well so that they are visited in a
    // cell-coherent order (for a self-query the two grids coincide)
    const bool self = (query == support) && Nq == Ns && qry_stride == sup_stride;
    GridSlot gs, gq;
    if (!take_grid_slot(W, B, Ns, &gs)) return DSIR_ERR_WORKSPACE;
    KnnGridBuildParams BP{};
    BP.pts4 = sup4; BP.pts_bs = Ns; BP.gmax = KNN_GRID_GMAX; BP.cells_per_point = GRID_CELLS_PER_POINT;
    BP.n[0] = Ns; BP.hdr[0] = gs.hdr; BP.cell_start[0] = gs.cell_start; BP.sorted[0] = gs.sorted;
    if ((rc = launch_knn_grid_build(BP, 1, B, st))) return rc;
    gq = gs;
    if (!self && Nq > 0) {
        float4 *q4 = W.take<float4>((size_t)B * Nq);
        if (!W.ok() || !take_grid_slot(W, B, Nq, &gq)) return DSIR_ERR_WORKSPACE;
        if ((rc = launch_pack_xyz4(query, qry_stride, (long long)B * Nq, q4, st))) return rc;
        KnnGridBuildParams QP = BP;
        QP.pts4 = q4; QP.pts_bs = Nq; QP.n[0] = Nq; QP.hdr[0] = gq.hdr; QP.cell_start[0] = gq.cell_start; QP.sorted[0] = gq.sorted;
        if ((rc = launch_knn_grid_build(QP, 1, B, st))) return rc;
    }
    KnnGridQueryParams Q{};
    Q.hdr = gs.hdr; Q.cell_start = gs.cell_start; Q.sorted = gs.sorted; Q.gmax = KNN_GRID_GMAX; Q.Ns = Ns;
    Q.q_sorted = gq.sorted; Q.q_sorted_bs = Nq;
    Q.Nq = Nq; Q.k = k; Q.r0_cells = GRID_R0_CELLS;
    Q.idx = idx; Q.idx_bs = (long long)Nq * k; Q.dist2 = dist2;
    return launch_knn_grid_query(Q, B, st);
}

static int pyramid_levels(int N, const int *ratios, int L, PyramidLevels *lv) {
    if (L < 1 || L > DSIR_MAX_LEVELS) return DSIR_ERR_UNSUPPORTED;
    lv->L = L;
    int n = N, off = 0, offsub = 0;
    for (int l = 0; l < L; ++l) {
        if (ratios[l] < 1) return DSIR_ERR_BAD_ARG;
        lv->n[l] = n; lv->m[l] = n / ratios[l]; lv->off[l] = off; lv->offsub[l] = offsub;
        off += n; offsub += n / ratios[l];
        n = n / ratios[l];
    }
    lv->sumN = off; lv->sumSub = offsub;
    return DSIR_OK;
}

// distinct support sizes of a pyramid that get a spatial structure (one type per pyramid, decided by the largest
// cloud): n[0], n[1], ..., and the last sub-cloud m[L-1]
static int pyramid_grid_sizes(const PyramidLevels &lv, int algo, int k, int *sizes, int *sp_out) {
    int ng = 0;
    const int sp = knn_structure(algo, lv.n[0] >= KNN_TREE_MIN_POINTS ? lv.n[0] : KNN_TREE_MIN_POINTS, lv.n[0], k, true);
    *sp_out = sp;
    if (sp == SP_BRUTE) return 0;
    for (int l = 0; l <= lv.L; ++l) {
        int n = l < lv.L ? lv.n[l] : lv.m[lv.L - 1];
        bool dup = false;
        for (int g = 0; g < ng; ++g) dup = dup || sizes[g] == n;
        const bool indexed = (algo == DSIR_KNN_GRID || algo == DSIR_KNN_TREE) ? true : n >= KNN_TREE_MIN_POINTS;
        if (!dup && n > 0 && indexed) sizes[ng++] = n;
    }
    return ng;
}

size_t dsir_knn_pyramid_workspace_bytes(int B, int N, int k, const int *ratios, int L, int algo) {
    if (B <= 0 || N <= 0) return 256;
    size_t bytes = ws_block((size_t)B * N * sizeof(float4)) + 256;
    PyramidLevels lv;
    if (ratios && pyramid_levels(N, ratios, L, &lv) == DSIR_OK) {
        int sizes[DSIR_MAX_LEVELS + 1], sp = SP_BRUTE;
        int ng = pyramid_grid_sizes(lv, algo, k, sizes, &sp);
        for (int g = 0; g < ng; ++g) bytes += sp == SP_TREE ? knn_tree_slot_bytes(B, sizes[g]) : grid_slot_bytes(B, sizes[g]);
    }
    return bytes;
}

int dsir_knn_pyramid(const float *pts, int pt_stride, int B, int N, const int *ratios, int L, int k, float *xyz_cat,
                     int64_t *neigh, int64_t *sub, int64_t *interp, void *ws, size_t ws_bytes, int algo,
                     dsir_stream_t stream) {
    if (!pts || !ratios || !neigh || !sub || !interp || B <= 0 || N <= 0 || pt_stride < 3 || k <= 0) return DSIR_ERR_BAD_ARG;
    if (algo < DSIR_KNN_AUTO || algo > DSIR_KNN_TREE) return DSIR_ERR_BAD_ARG;
    if (k > 32) return DSIR_ERR_UNSUPPORTED;
    if (algo == DSIR_KNN_TREE && N > KNN_TREE_MAX_POINTS) return DSIR_ERR_UNSUPPORTED;
    PyramidLevels lv;
    int rc = pyramid_levels(N, ratios, L, &lv);
    if (rc) return rc;
    for (int l = 0; l < L; ++l)
        if (lv.n[l] < k || lv.m[l] < 1) return DSIR_ERR_KNN_TOO_FEW;
    cudaStream_t st = (cudaStream_t)stream;
    Workspace W(ws, ws_bytes);
    float4 *pts4 = W.take<float4>((size_t)B * N);
    if (!W.ok()) return DSIR_ERR_WORKSPACE;
    if ((rc = launch_pack_xyz4(pts, pt_stride, (long long)B * N, pts4, st))) return rc;
    if (xyz_cat && (rc = launch_pyramid_xyz(pts, pt_stride, B, N, lv, xyz_cat, st))) return rc;

    // one build launch for every grid of the pyramid (every level cloud is a prefix of the packed cloud)
    int sizes[DSIR_MAX_LEVELS + 1], sp = SP_BRUTE;
    const int ng = pyramid_grid_sizes(lv, algo, k, sizes, &sp);
    GridSlot slots[DSIR_MAX_LEVELS + 1];
    KnnTreeView trees[DSIR_MAX_LEVELS + 1];
    if (ng > 0 && sp == SP_TREE) {
        for (int g = 0; g < ng; ++g)
            if (!knn_tree_take_slot(W, B, sizes[g], &trees[g])) return DSIR_ERR_WORKSPACE;
        if ((rc = launch_knn_tree_build(pts4, N, trees, ng, B, st))) return rc;
    } else if (ng > 0) {
        KnnGridBuildParams BP{};
        BP.pts4 = pts4; BP.pts_bs = N; BP.gmax = KNN_GRID_GMAX; BP.cells_per_point = GRID_CELLS_PER_POINT;
        for (int g = 0; g < ng; ++g) {
            if (!take_grid_slot(W, B, sizes[g], &slots[g])) return DSIR_ERR_WORKSPACE;
            BP.n[g] = sizes[g]; BP.hdr[g] = slots[g].hdr; BP.cell_start[g] = slots[g].cell_start; BP.sorted[g] = slots[g].sorted;
        }
        if ((rc = launch_knn_grid_build(BP, ng, B, st))) return rc;
    }
    auto find_slot = [&](int n) -> const GridSlot * {
        if (sp != SP_GRID) return nullptr;
        for (int g = 0; g < ng; ++g)
            if (slots[g].n == n) return &slots[g];
        return nullptr;
    };
    auto find_tree = [&](int n) -> const KnnTreeView * {
        if (sp != SP_TREE) return nullptr;
        for (int g = 0; g < ng; ++g)
            if (trees[g].n == n) return &trees[g];
        return nullptr;
    };

    // fork: level-0 self-kNN stays on `st`; coarser self-kNNs -> aux stream A, every 1-NN up-sampling -> aux stream B
    int dev = 0;
    cudaGetDevice(&dev);
    const bool fork = KNN_FORK && L > 1 && g_aux.ready(dev);
    cudaStream_t st_main = st, st_self = st, st_up = st;
    if (fork) {
        cudaEvent_t ev = g_aux.event();
        st_self = g_aux.stream();
        st_up = g_aux.stream();
        DSIR_CUDA_TRY(cudaEventRecord(ev, st_main));
        DSIR_CUDA_TRY(cudaStreamWaitEvent(st_self, ev, 0));
        DSIR_CUDA_TRY(cudaStreamWaitEvent(st_up, ev, 0));
    }
    auto join = [&]() -> int {
        if (!fork) return DSIR_OK;
        cudaEvent_t e1 = g_aux.event(), e2 = g_aux.event();
        DSIR_CUDA_TRY(cudaEventRecord(e1, st_self));
        DSIR_CUDA_TRY(cudaEventRecord(e2, st_up));
        DSIR_CUDA_TRY(cudaStreamWaitEvent(st_main, e1, 0));
        DSIR_CUDA_TRY(cudaStreamWaitEvent(st_main, e2, 0));
        return DSIR_OK;
    };

    for (int l = 0; l < L; ++l) {
        const GridSlot *gl = find_slot(lv.n[l]);   // grid over the level cloud (support of the self-kNN, query order)
        const GridSlot *gm = find_slot(lv.m[l]);   // grid over the sub-cloud (support of the 1-NN up-sampling)
        const KnnTreeView *tl = find_tree(lv.n[l]), *tm = find_tree(lv.m[l]);
        st = l == 0 ? st_main : st_self;
        int64_t *nb = neigh + (size_t)lv.off[l] * k;
        int64_t *pool = sub + (size_t)lv.offsub[l] * k;
        int64_t *up = interp + lv.off[l];
        // ---- self-kNN of the level cloud (data_base.py:165); rows < m[l] are also the pooling indices (:168) ----
        if (tl) {
            KnnTreeQueryParams Q{};
            Q.sup = *tl; Q.qry = *tl; Q.self = 1; Q.k = k;
            Q.idx = nb; Q.idx_bs = (long long)lv.sumN * k;
            Q.idx2 = pool; Q.idx2_bs = (long long)lv.sumSub * k; Q.idx2_rows = lv.m[l];
            if ((rc = launch_knn_tree_query(Q, B, st))) { join(); return rc; }
        } else if (gl) {
            KnnGridQueryParams Q{};
            Q.hdr = gl->hdr; Q.cell_start = gl->cell_start; Q.sorted = gl->sorted; Q.gmax = KNN_GRID_GMAX; Q.Ns = lv.n[l];
            Q.q_sorted = gl->sorted; Q.q_sorted_bs = lv.n[l];
            Q.Nq = lv.n[l]; Q.k = k; Q.r0_cells = GRID_R0_CELLS;
            Q.idx = nb; Q.idx_bs = (long long)lv.sumN * k;
            Q.idx2 = pool; Q.idx2_bs = (long long)lv.sumSub * k; Q.idx2_rows = lv.m[l];
            if ((rc = launch_knn_grid_query(Q, B, st))) { join(); return rc; }
        } else {
            KnnBruteParams P{};
            P.sup4 = pts4; P.sup_bs = N;
            P.query = (const float *)pts4; P.qry_bs = (long long)N * 4; P.qry_stride = 4;
            P.Ns = lv.n[l]; P.Nq = lv.n[l]; P.k = k;
            P.idx = nb; P.idx_bs = (long long)lv.sumN * k;
            P.idx2 = pool; P.idx2_bs = (long long)lv.sumSub * k; P.idx2_rows = lv.m[l];
            if ((rc = launch_knn_brute(P, B, st))) { join(); return rc; }
        }
        // ---- 1-NN of every level point into the sub-cloud (data_base.py:170) ----
        st = st_up;
        if (tm && tl) {
            KnnTreeQueryParams Q{};
            Q.sup = *tm; Q.qry = *tl; Q.self = 0; Q.k = 1;
            Q.idx = up; Q.idx_bs = lv.sumN;
            if ((rc = launch_knn_tree_query(Q, B, st))) { join(); return rc; }
        } else if (gm) {
            KnnGridQueryParams Q{};
            Q.hdr = gm->hdr; Q.cell_start = gm->cell_start; Q.sorted = gm->sorted; Q.gmax = KNN_GRID_GMAX; Q.Ns = lv.m[l];
            if (gl) { Q.q_sorted = gl->sorted; Q.q_sorted_bs = lv.n[l]; }
            else { Q.query = (const float *)pts4; Q.qry_bs = (long long)N * 4; Q.qry_stride = 4; }
            Q.Nq = lv.n[l]; Q.k = 1; Q.r0_cells = GRID_R0_CELLS;
            Q.idx = up; Q.idx_bs = lv.sumN;
            if ((rc = launch_knn_grid_query(Q, B, st))) { join(); return rc; }
        } else {
            KnnBruteParams U{};
            U.sup4 = pts4; U.sup_bs = N;
            U.query = (const float *)pts4; U.qry_bs = (long long)N * 4; U.qry_stride = 4;
            U.Ns = lv.m[l]; U.Nq = lv.n[l]; U.k = 1;
            U.idx = up; U.idx_bs = lv.sumN;
            if ((rc = launch_knn_brute(U, B, st))) { join(); return rc; }
        }
    }
    return join();
}

/* ------------------------------------------------------------------ match ---------------------- */
static bool feat_ok(const dsir_feat &f) { return f.ptr != nullptr; }

size_t dsir_match_dense_workspace_bytes(int B, int J, int K) {
    if (B <= 0 || J <= 0 || K <= 0) return 256;
    return ws_block((size_t)B * J * sizeof(float)) + ws_block((size_t)B * K * sizeof(float)) + 256;
}

int dsir_match_dense(dsir_feat fs, dsir_feat fr, int B, int C, int J, int K, int metric, float *dist, void *ws,
                     size_t ws_bytes, dsir_stream_t stream) {
    if (!feat_ok(fs) || !feat_ok(fr) || !dist || B <= 0 || C <= 0 || J <= 0 || K <= 0) return DSIR_ERR_BAD_ARG;
    if (metric < DSIR_METRIC_L2 || metric > DSIR_METRIC_SQDIFF_SQRT) return DSIR_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    MatchParams P{};
    P.fs = fs; P.fr = fr; P.B = B; P.C = C; P.J = J; P.K = K;
    P.dense = dist; P.metric = metric;
    if (metric == DSIR_METRIC_L2 || metric == DSIR_METRIC_EUCLIDEAN) {
        Workspace W(ws, ws_bytes);
        float *ns = W.take<float>((size_t)B * J);
        float *nr = W.take<float>((size_t)B * K);
        if (!W.ok()) return DSIR_ERR_WORKSPACE;
        int rc;
        if ((rc = launch_sqnorm(fs, B, C, J, ns, st))) return rc;
        if ((rc = launch_sqnorm(fr, B, C, K, nr, st))) return rc;
        P.ns = ns; P.nr = nr;
    }
    return launch_match_fp32(P, MATCH_MODE_DENSE, st);
}

size_t dsir_match_argmin_workspace_bytes(int B, int C, int J, int K, int algo) {
    if (B <= 0 || J <= 0 || K <= 0) return 256;
    size_t bytes = ws_block((size_t)B * J * sizeof(float)) + ws_block((size_t)B * K * sizeof(float)) + 256;
    if (algo != DSIR_MATCH_FP32) bytes += match_tc_workspace_bytes(B, C, J, K);
    return bytes;
}

}  // extern "C"

// reuse_prep: iterations 2.. of the alignment loop (same features, same workspace): norms / operand copies are kept
static int match_argmin_impl(dsir_feat fs, dsir_feat fr, int B, int C, int J, int K, int64_t *idx, float *min_d, void *ws,
                             size_t ws_bytes, int algo, bool reuse_prep, cudaStream_t st, const int64_t *prior = nullptr) {
    if (!feat_ok(fs) || !feat_ok(fr) || !idx || B <= 0 || C <= 0 || J <= 0 || K <= 0) return DSIR_ERR_BAD_ARG;
    Workspace W(ws, ws_bytes);
    float *ns = W.take<float>((size_t)B * J);
    float *nr = W.take<float>((size_t)B * K);
    if (!W.ok()) return DSIR_ERR_WORKSPACE;
    int rc;
    MatchParams P{};
    P.fs = fs; P.fr = fr; P.B = B; P.C = C; P.J = J; P.K = K; P.ns = ns; P.nr = nr;
    P.idx = idx; P.min_d = min_d; P.reuse_prep = reuse_prep ? 1 : 0; P.prior_idx = prior;
    bool tc_ok = match_tc_supported(fs, fr, B, C, J, K);
    if (algo == DSIR_MATCH_TC && !tc_ok) return DSIR_ERR_UNSUPPORTED;
    if (algo == DSIR_MATCH_TC || (algo == DSIR_MATCH_AUTO && tc_ok && match_tc_profitable(B, C, J, K))) {
        size_t used = W.off;
        return launch_match_tc(P, (char *)ws + used, ws_bytes - used, st);   // fills ns / nr itself (norms + maxima fused)
    }
    if (!reuse_prep) {
        if ((rc = launch_sqnorm(fs, B, C, J, ns, st))) return rc;
        if ((rc = launch_sqnorm(fr, B, C, K, nr, st))) return rc;
    }
    return launch_match_fp32(P, MATCH_MODE_ARGMIN, st);
}

extern "C" {

int dsir_match_argmin(dsir_feat fs, dsir_feat fr, int B, int C, int J, int K, int64_t *idx, float *min_d, void *ws,
                      size_t ws_bytes, int algo, dsir_stream_t stream) {
    return match_argmin_impl(fs, fr, B, C, J, K, idx, min_d, ws, ws_bytes, algo, false, (cudaStream_t)stream);
}

int dsir_match_argmin_hint(dsir_feat fs, dsir_feat fr, int B, int C, int J, int K, int64_t *idx, float *min_d,
                           const int64_t *prior_idx, void *ws, size_t ws_bytes, int algo, dsir_stream_t stream) {
    return match_argmin_impl(fs, fr, B, C, J, K, idx, min_d, ws, ws_bytes, algo, false, (cudaStream_t)stream, prior_idx);
}

int dsir_match_argmin_rescued_rows(const void *ws, size_t ws_bytes, int B, int C, int J, int K, int32_t *host_out,
                                   dsir_stream_t stream) {
    if (!ws || !host_out || B <= 0 || C <= 0 || J <= 0 || K <= 0) return DSIR_ERR_BAD_ARG;
    size_t used = ws_block((size_t)B * J * sizeof(float)) + ws_block((size_t)B * K * sizeof(float));
    if (used >= ws_bytes) return DSIR_ERR_WORKSPACE;
    return match_tc_rescued_rows((const char *)ws + used, B, C, J, K, host_out, (cudaStream_t)stream);
}

int dsir_match_argmin_filter_timing(const void *ws, size_t ws_bytes, int B, int C, int J, int K, double *host_out,
                                    dsir_stream_t stream) {
    if (!ws || !host_out || B <= 0 || C <= 0 || J <= 0 || K <= 0) return DSIR_ERR_BAD_ARG;
    size_t used = ws_block((size_t)B * J * sizeof(float)) + ws_block((size_t)B * K * sizeof(float));
    if (used >= ws_bytes) return DSIR_ERR_WORKSPACE;
    return match_tc_filter_timing((const char *)ws + used, B, C, J, K, host_out, (cudaStream_t)stream);
}

int dsir_match_argmin_filter_trace(const void *ws, size_t ws_bytes, int B, int C, int J, int K, uint32_t *host_out,
                                   dsir_stream_t stream) {
    if (!ws || !host_out || B <= 0 || C <= 0 || J <= 0 || K <= 0) return DSIR_ERR_BAD_ARG;
    size_t used = ws_block((size_t)B * J * sizeof(float)) + ws_block((size_t)B * K * sizeof(float));
    if (used >= ws_bytes) return DSIR_ERR_WORKSPACE;
    return match_tc_filter_trace((const char *)ws + used, B, C, J, K, host_out, (cudaStream_t)stream);
}

size_t dsir_match_soft_workspace_bytes(int B, int C, int J, int K) {
    if (B <= 0 || J <= 0 || K <= 0) return 256;
    size_t fp32 = ws_block((size_t)B * J * sizeof(float)) + ws_block((size_t)B * K * sizeof(float)) + 256;
    if (match_tc_soft_supported(B, C, J, K)) {
        size_t tc = match_tc_soft_workspace_bytes(B, C, J, K);
        return tc > fp32 ? tc : fp32;
    }
    return fp32;
}

// Top-k soft weights.  Fused route (match_tc.cu, launch_match_tc_topk): two tensor-core sweeps + exact re-scoring of the
// listed columns, nothing of size J x K.  Materialising route (small problems, C > 64, a column bias - the order is then not
// the distance order): row chunks of the exact fp32 distance matrix (<= 256 MB, or whatever the workspace holds) + one warp
// per row.  Both give the same bits.
static size_t soft_topk_chunk_rows(int B, int J, int K, size_t budget = (size_t)256 << 20) {
    long long rows = (long long)(budget / ((size_t)B * K * 4 + (size_t)B * 4));
    if (rows < 1) rows = 1;
    if (rows > J) rows = J;
    return (size_t)rows;
}
static size_t soft_topk_chunk_fixed(int B, int J, int K) { return ws_block((size_t)B * K * 4) + ws_block((size_t)B * J * 4) + 1024; }
size_t dsir_match_soft_topk_workspace_bytes(int B, int C, int J, int K, int topk) {
    size_t base = dsir_match_soft_workspace_bytes(B, C, J, K);
    if (topk <= 0 || B <= 0 || J <= 0 || K <= 0) return base;
    if (match_tc_topk_supported(B, C, J, K, topk)) {
        const size_t fused = ws_block((size_t)B * J * 4) + match_tc_topk_workspace_bytes(B, C, J, K, topk) + 512;
        const size_t one_row = soft_topk_chunk_fixed(B, J, K) + ws_block((size_t)B * K * 4) + ws_block((size_t)B * 4);
        return base + (fused > one_row ? fused : one_row);
    }
    const size_t Jc = soft_topk_chunk_rows(B, J, K);
    return base + soft_topk_chunk_fixed(B, J, K) + ws_block((size_t)B * Jc * K * 4) + ws_block((size_t)B * Jc * 4);
}

static int match_soft_core(dsir_feat fs, dsir_feat fr, int B, int C, int J, int K, const float *beta, const float *alpha,
                           const float *col_bias, const float *xyz_ref, float *y_soft, float *lse, void *ws, size_t ws_bytes,
                           cudaStream_t st, bool reuse_prep = false) {
    if (match_tc_soft_supported(B, C, J, K) && (y_soft || lse)) {   // tcgen05: fp16 x2 split + online softmax in the TMEM epilogue
        MatchParams T{};
        T.fs = fs; T.fr = fr; T.B = B; T.C = C; T.J = J; T.K = K;
        T.beta = beta; T.alpha = alpha; T.col_bias = col_bias; T.xyz_ref = xyz_ref; T.y_soft = y_soft; T.lse = lse;
        T.reuse_prep = reuse_prep ? 1 : 0;
        return launch_match_tc_soft(T, ws, ws_bytes, st);
    }
    Workspace W(ws, ws_bytes);
    float *ns = W.take<float>((size_t)B * J);
    float *nr = W.take<float>((size_t)B * K);
    if (!W.ok()) return DSIR_ERR_WORKSPACE;
    int rc;
    if (!reuse_prep) {
        if ((rc = launch_sqnorm(fs, B, C, J, ns, st))) return rc;
        if ((rc = launch_sqnorm(fr, B, C, K, nr, st))) return rc;
    }
    MatchParams P{};
    P.fs = fs; P.fr = fr; P.B = B; P.C = C; P.J = J; P.K = K; P.ns = ns; P.nr = nr;
    P.beta = beta; P.alpha = alpha; P.col_bias = col_bias; P.xyz_ref = xyz_ref; P.y_soft = y_soft; P.lse = lse;
    return launch_match_fp32(P, MATCH_MODE_SOFT, st);
}

int dsir_match_soft(dsir_feat fs, dsir_feat fr, int B, int C, int J, int K, const float *beta, const float *alpha,
                    const float *col_bias, const float *xyz_ref, float *y_soft, float *lse, int topk,
                    int64_t *topk_idx, float *topk_w, void *ws, size_t ws_bytes, dsir_stream_t stream) {
    if (!feat_ok(fs) || !feat_ok(fr) || !beta || !alpha || B <= 0 || C <= 0 || J <= 0 || K <= 0) return DSIR_ERR_BAD_ARG;
    if (y_soft && !xyz_ref) return DSIR_ERR_BAD_ARG;
    if (topk < 0 || (topk > 0 && (!topk_idx || !topk_w))) return DSIR_ERR_BAD_ARG;
    if (topk > 32 || topk > K) return topk > K ? DSIR_ERR_BAD_ARG : DSIR_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (topk == 0) return match_soft_core(fs, fr, B, C, J, K, beta, alpha, col_bias, xyz_ref, y_soft, lse, ws, ws_bytes, st);
    const size_t base = dsir_match_soft_workspace_bytes(B, C, J, K);
    if (!ws || ws_bytes < dsir_match_soft_topk_workspace_bytes(B, C, J, K, topk)) return DSIR_ERR_WORKSPACE;
    int rc;
    if (match_tc_topk_supported(B, C, J, K, topk) && col_bias == nullptr) {
        // ---- fused: lse (and y) from the online-softmax pass, the k best columns from two filter sweeps + exact re-scoring
        Workspace W((char *)ws + base, ws_bytes - base);
        float *lse_tmp = W.take<float>((size_t)B * J);
        if (!W.ok()) return DSIR_ERR_WORKSPACE;
        float *lse_use = lse ? lse : lse_tmp;
        if ((rc = match_soft_core(fs, fr, B, C, J, K, beta, alpha, col_bias, xyz_ref, y_soft, lse_use, ws, base, st))) return rc;
        MatchParams T{};
        T.fs = fs; T.fr = fr; T.B = B; T.C = C; T.J = J; T.K = K; T.beta = beta; T.alpha = alpha; T.lse = lse_use;
        const size_t used = base + ws_block((size_t)B * J * 4);
        return launch_match_tc_topk(T, topk, topk_idx, topk_w, (char *)ws + used, ws_bytes - used, st);
    }
    // ---- materialising: the k largest weights of a row come from row chunks of the exact fp32 distance matrix (DENSE
    //      kernel), one warp per row ----
    Workspace W((char *)ws + base, ws_bytes - base);
    float *nr = W.take<float>((size_t)B * K);
    float *lse_tmp = W.take<float>((size_t)B * J);
    if (!W.ok()) return DSIR_ERR_WORKSPACE;
    const size_t left = ws_bytes - base - soft_topk_chunk_fixed(B, J, K);
    int Jc = (int)soft_topk_chunk_rows(B, J, K, left < ((size_t)256 << 20) ? left : ((size_t)256 << 20));
    while (Jc > 1 && ws_block((size_t)B * Jc * K * 4) + ws_block((size_t)B * Jc * 4) > left) --Jc;
    float *chunk = W.take<float>((size_t)B * Jc * K);
    float *ns_c = W.take<float>((size_t)B * Jc);
    if (!W.ok()) return DSIR_ERR_WORKSPACE;
    float *lse_use = lse ? lse : lse_tmp;
    if ((rc = match_soft_core(fs, fr, B, C, J, K, beta, alpha, col_bias, xyz_ref, y_soft, lse_use, ws, base, st))) return rc;
    if ((rc = launch_sqnorm(fr, B, C, K, nr, st))) return rc;
    for (int j0 = 0; j0 < J; j0 += Jc) {
        const int jc = J - j0 < Jc ? J - j0 : Jc;
        dsir_feat fsc = fs;
        fsc.ptr = fs.ptr + (size_t)j0 * fs.point_stride;
        if ((rc = launch_sqnorm(fsc, B, C, jc, ns_c, st))) return rc;
        MatchParams D{};
        D.fs = fsc; D.fr = fr; D.B = B; D.C = C; D.J = jc; D.K = K; D.ns = ns_c; D.nr = nr; D.dense = chunk; D.metric = DSIR_METRIC_L2;
        if ((rc = launch_match_fp32(D, MATCH_MODE_DENSE, st))) return rc;
        if ((rc = launch_row_topk(chunk, B, jc, K, beta, alpha, col_bias, lse_use, J, j0, topk, topk_idx, topk_w, (long long)J * topk, st))) return rc;
    }
    return DSIR_OK;
}

int dsir_match_soft_topk_fused(int B, int C, int J, int K, int topk) {
    return (B > 0 && J > 0 && K > 0 && match_tc_topk_supported(B, C, J, K, topk)) ? 1 : 0;
}

int dsir_match_soft_topk_exhaustive_rows(const void *ws, size_t ws_bytes, int B, int C, int J, int K, int topk, int32_t *host_out,
                                         dsir_stream_t stream) {
    if (!ws || !host_out || B <= 0 || J <= 0 || K <= 0) return DSIR_ERR_BAD_ARG;
    if (!match_tc_topk_supported(B, C, J, K, topk)) return DSIR_ERR_UNSUPPORTED;
    if (ws_bytes < dsir_match_soft_topk_workspace_bytes(B, C, J, K, topk)) return DSIR_ERR_WORKSPACE;
    const size_t used = dsir_match_soft_workspace_bytes(B, C, J, K) + ws_block((size_t)B * J * 4);
    return match_tc_topk_exhaustive_rows((const char *)ws + used, B, C, J, K, topk, host_out, (cudaStream_t)stream);
}

int dsir_match_soft_sweep(dsir_feat fs, dsir_feat fr, int B, int C, int J, int K, const float *beta, const float *alpha,
                          const float *col_bias, const float *xyz_ref, float *y_soft, float *lse, int reuse_prep, void *ws,
                          size_t ws_bytes, dsir_stream_t stream) {
    if (!feat_ok(fs) || !feat_ok(fr) || !beta || !alpha || B <= 0 || C <= 0 || J <= 0 || K <= 0) return DSIR_ERR_BAD_ARG;
    if (y_soft && !xyz_ref) return DSIR_ERR_BAD_ARG;
    return match_soft_core(fs, fr, B, C, J, K, beta, alpha, col_bias, xyz_ref, y_soft, lse, ws, ws_bytes, (cudaStream_t)stream,
                           reuse_prep != 0);
}

int dsir_gather_points(const float *in, int B, int C, int N, const int64_t *idx, int M, float *out,
                       dsir_stream_t stream) {
    if (!in || !idx || !out || B <= 0 || C <= 0 || N <= 0 || M < 0) return DSIR_ERR_BAD_ARG;
    return launch_gather_points(in, B, C, N, idx, M, out, (cudaStream_t)stream);
}

/* ------------------------------------------------------------------ KNN consumers / Sinkhorn --- */
int dsir_gather_neighbours(const float *in, int B, int C, int N, const int64_t *idx, int M, int k, float *out,
                           dsir_stream_t stream) {
    if (!in || !idx || !out || B <= 0 || C <= 0 || N <= 0 || M < 0 || k <= 0) return DSIR_ERR_BAD_ARG;
    return launch_gather_neighbours(in, B, C, N, idx, M, k, out, (cudaStream_t)stream);
}

int dsir_rel_pos_encoding(const float *xyz, int B, int N, const int64_t *idx, int k, float *out, dsir_stream_t stream) {
    if (!xyz || !idx || !out || B <= 0 || N <= 0 || k <= 0) return DSIR_ERR_BAD_ARG;
    return launch_rel_pos_encoding(xyz, B, N, idx, k, out, (cudaStream_t)stream);
}

int dsir_pool_max(const float *in, int B, int C, int N, const int64_t *idx, int M, int k, float *out, dsir_stream_t stream) {
    if (!in || !idx || !out || B <= 0 || C <= 0 || N <= 0 || M < 0 || k <= 0) return DSIR_ERR_BAD_ARG;
    return launch_pool_max(in, B, C, N, idx, M, k, out, (cudaStream_t)stream);
}

size_t dsir_sinkhorn_workspace_bytes(int B, int J, int K) {
    if (B <= 0 || J <= 0 || K <= 0) return 256;
    return ws_block((size_t)B * J * sizeof(float)) + ws_block((size_t)B * K * sizeof(float)) + 256;
}

size_t dsir_log_optimal_transport_workspace_bytes(int B, int M, int N) {
    if (B <= 0 || M <= 0 || N <= 0) return 256;
    return ws_block((size_t)B * (M + 1) * sizeof(float)) + ws_block((size_t)B * (N + 1) * sizeof(float)) + 256;
}

int dsir_log_optimal_transport(const float *scores, int B, int M, int N, const float *alpha, int iters, float *out, void *ws,
                               size_t ws_bytes, dsir_stream_t stream) {
    if (!scores || !alpha || !out || B <= 0 || M <= 0 || N <= 0 || iters < 0) return DSIR_ERR_BAD_ARG;
    Workspace W(ws, ws_bytes);
    float *u = W.take<float>((size_t)B * (M + 1));
    float *v = W.take<float>((size_t)B * (N + 1));
    if (!W.ok()) return DSIR_ERR_WORKSPACE;
    return launch_log_ot(scores, B, M, N, alpha, iters, out, u, v, (cudaStream_t)stream);
}

int dsir_sinkhorn(const float *log_alpha, int B, int J, int K, int n_iters, int slack, float *out, void *ws, size_t ws_bytes,
                  dsir_stream_t stream) {
    if (!log_alpha || !out || B <= 0 || J <= 0 || K <= 0 || n_iters < 0) return DSIR_ERR_BAD_ARG;
    Workspace W(ws, ws_bytes);
    float *u = W.take<float>((size_t)B * J);
    float *v = W.take<float>((size_t)B * K);
    if (!W.ok()) return DSIR_ERR_WORKSPACE;
    return launch_sinkhorn(log_alpha, B, J, K, n_iters, slack ? 1 : 0, out, u, v, (cudaStream_t)stream);
}

/* ------------------------------------------------------------------ Kabsch --------------------- */
size_t dsir_kabsch_workspace_bytes(int B, int M) {
    if (B <= 0 || M <= 0) return 256;
    return ws_block((size_t)B * kabsch_num_blocks(M) * KB_NMOM * sizeof(double)) + 256;
}

static int kabsch_common(dsir_points src, dsir_points tgt, const float *w, int64_t w_bs, const int64_t *gather, int B,
                         int M, int n_tgt, double **partials, int *nblk, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (!src.ptr || !tgt.ptr || B <= 0 || M <= 0) return DSIR_ERR_BAD_ARG;
    Workspace W(ws, ws_bytes);
    *nblk = kabsch_num_blocks(M);
    *partials = W.take<double>((size_t)B * (*nblk) * KB_NMOM);
    if (!W.ok()) return DSIR_ERR_WORKSPACE;
    KabschParams P{};
    P.src = src; P.tgt = tgt; P.w = w; P.w_bs = w_bs; P.gather = gather; P.B = B; P.M = M; P.n_tgt = n_tgt; P.partials = *partials;
    return launch_kabsch_moments(P, *nblk, st);
}

int dsir_kabsch(dsir_points src, dsir_points tgt, const float *w, int64_t w_batch_stride, const int64_t *gather,
                int B, int M, int n_tgt, float *T, int32_t *status, double *moments, void *ws, size_t ws_bytes,
                dsir_stream_t stream) {
    if (!T) return DSIR_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    double *partials; int nblk;
    int rc = kabsch_common(src, tgt, w, w_batch_stride, gather, B, M, n_tgt, &partials, &nblk, ws, ws_bytes, st);
    if (rc) return rc;
    return launch_kabsch_solve(partials, nblk, B, T, status, moments, nullptr, nullptr, 0, st);
}

int dsir_kabsch_moments(dsir_points src, dsir_points tgt, const float *w, int64_t w_batch_stride,
                        const int64_t *gather, int B, int M, int n_tgt, double *moments, void *ws, size_t ws_bytes,
                        dsir_stream_t stream) {
    if (!moments) return DSIR_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    double *partials; int nblk;
    int rc = kabsch_common(src, tgt, w, w_batch_stride, gather, B, M, n_tgt, &partials, &nblk, ws, ws_bytes, st);
    if (rc) return rc;
    return launch_kabsch_reduce(partials, nblk, B, moments, st);
}

int dsir_kabsch_from_moments(const double *moments, int B, float *T, int32_t *status, dsir_stream_t stream) {
    if (!moments || !T || B <= 0) return DSIR_ERR_BAD_ARG;
    return launch_kabsch_solve(moments, 1, B, T, status, nullptr, nullptr, nullptr, 0, (cudaStream_t)stream);
}

int dsir_soft_targets(const float *weights, int64_t w_batch_stride, int64_t w_row_stride, const float *tgt, int B, int M, int N,
                      float *y_soft, float *rowmass, dsir_stream_t stream) {
    if (!weights || !tgt || !y_soft || !rowmass || B <= 0 || M <= 0 || N <= 0) return DSIR_ERR_BAD_ARG;
    return launch_soft_targets(weights, w_batch_stride, w_row_stride, tgt, B, M, N, y_soft, rowmass, (cudaStream_t)stream);
}

int dsir_kabsch_soft(dsir_points src, const float *y_soft, const float *rowmass, int B, int M, float *T,
                     int32_t *status, void *ws, size_t ws_bytes, dsir_stream_t stream) {
    if (!y_soft || !rowmass || !T) return DSIR_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    dsir_points tgt{y_soft, (int64_t)M * 3, 3, 1};
    double *partials; int nblk;
    int rc = kabsch_common(src, tgt, rowmass, M, nullptr, B, M, 0, &partials, &nblk, ws, ws_bytes, st);
    if (rc) return rc;
    return launch_kabsch_solve(partials, nblk, B, T, status, nullptr, nullptr, nullptr, 1, st);
}

/* ------------------------------------------------------------------ SE(3) ---------------------- */
int dsir_se3_apply(const float *T, int64_t T_batch_stride, dsir_points pts, int B, int N, float *out,
                   int64_t out_batch_stride, int64_t out_point_stride, int64_t out_coord_stride,
                   int rotate_only, dsir_stream_t stream) {
    if (!T || !pts.ptr || !out || B <= 0 || N < 0) return DSIR_ERR_BAD_ARG;
    return launch_se3_apply(T, T_batch_stride, pts, B, N, out, out_batch_stride, out_point_stride, out_coord_stride,
                            rotate_only, (cudaStream_t)stream);
}

int dsir_se3_compose(const float *a, int64_t a_bs, const float *b, int64_t b_bs, int B, float *out,
                     dsir_stream_t stream) {
    if (!a || !b || !out || B <= 0) return DSIR_ERR_BAD_ARG;
    return launch_se3_compose(a, a_bs, b, b_bs, B, out, (cudaStream_t)stream);
}

int dsir_se3_inverse(const float *T, int64_t T_bs, int B, float *out, dsir_stream_t stream) {
    if (!T || !out || B <= 0) return DSIR_ERR_BAD_ARG;
    return launch_se3_inverse(T, T_bs, B, out, (cudaStream_t)stream);
}

/* ------------------------------------------------------------------ loop ----------------------- */
size_t dsir_align_loop_workspace_bytes(int B, int C, int J, int K, int algo) {
    if (B <= 0 || J <= 0 || K <= 0) return 256;
    return dsir_match_argmin_workspace_bytes(B, C, J, K, algo) + dsir_kabsch_workspace_bytes(B, J) +
           ws_block((size_t)B * J * sizeof(int64_t)) + ws_block((size_t)B * 12 * sizeof(float)) + 256;
}

int dsir_align_loop(dsir_feat fs, dsir_feat fr, int B, int C, int J, int K, float *xyz_src, const float *xyz_ref,
                    const float *weights, int iters, float *transforms, int64_t *pred_idx, int32_t *status,
                    void *ws, size_t ws_bytes, int algo, dsir_stream_t stream) {
    if (!feat_ok(fs) || !feat_ok(fr) || !xyz_src || !xyz_ref || !transforms || B <= 0 || C <= 0 || J <= 0 || K <= 0 ||
        iters <= 0)
        return DSIR_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    Workspace W(ws, ws_bytes);
    int64_t *idx_scratch = W.take<int64_t>((size_t)B * J);
    float *T_it = W.take<float>((size_t)B * 12);
    if (!W.ok()) return DSIR_ERR_WORKSPACE;
    size_t match_bytes = dsir_match_argmin_workspace_bytes(B, C, J, K, algo);
    size_t kab_bytes = dsir_kabsch_workspace_bytes(B, J);
    if (W.off + match_bytes + kab_bytes > ws_bytes) return DSIR_ERR_WORKSPACE;
    char *match_ws = (char *)ws + W.off;
    char *kab_ws = match_ws + match_bytes;

    dsir_points src{xyz_src, (int64_t)3 * J, 1, J};  // [B,3,J]
    dsir_points ref{xyz_ref, (int64_t)3 * K, 1, K};  // [B,3,K]
    for (int it = 0; it < iters; ++it) {
        int64_t *idx = pred_idx ? pred_idx + (size_t)it * B * J : idx_scratch;
        // the features do not change inside this entry point: norms and operand copies are prepared once
        // ... and the previous iteration's correspondences prime the filter (a hint: results do not depend on it)
        const int64_t *prior = it > 0 ? (pred_idx ? pred_idx + (size_t)(it - 1) * B * J : idx_scratch) : nullptr;
        int rc = match_argmin_impl(fs, fr, B, C, J, K, idx, nullptr, match_ws, match_bytes, algo, it > 0, st, prior);  // :558-569
        if (rc) return rc;
        double *partials; int nblk;
        rc = kabsch_common(src, ref, weights, J, idx, B, J, K, &partials, &nblk, kab_ws, kab_bytes, st);    // :571,:588
        if (rc) return rc;
        float *T_cum = transforms + (size_t)it * B * 12;
        const float *T_prev = it > 0 ? transforms + (size_t)(it - 1) * B * 12 : nullptr;
        rc = launch_kabsch_solve(partials, nblk, B, T_it, status ? status + (size_t)it * B : nullptr, nullptr, T_prev,
                                 T_cum, 0, st);                                                              // :595
        if (rc) return rc;
        rc = launch_se3_apply(T_it, 12, src, B, J, xyz_src, (long long)3 * J, 1, J, 0, st);                  // :590
        if (rc) return rc;
    }
    return DSIR_OK;
}

/* ------------------------------------------------------------------ key points, evaluation ---- */
size_t dsir_keypoint_score_workspace_bytes(int B, int C, int N) {
    if (B <= 0 || C <= 0 || N <= 0) return 0;
    return keypoint_score_workspace_bytes(B, C, N);
}

int dsir_keypoint_score(const float *feat, const float *xyz, const float *prob, const int64_t *label,
                        const float *label_weights, int num_class, const int64_t *neigh_idx, int idx_stride, int k,
                        float ball_r, int B, int C, int N, float *score, void *ws, size_t ws_bytes, dsir_stream_t stream) {
    if (!feat || !xyz || !neigh_idx || !score || B <= 0 || C <= 0 || N <= 0 || k <= 0 || idx_stride < k) return DSIR_ERR_BAD_ARG;
    if (label && (!label_weights || num_class <= 0)) return DSIR_ERR_BAD_ARG;
    if (k > 32) return DSIR_ERR_UNSUPPORTED;
    return launch_keypoint_score(feat, xyz, prob, label, label_weights, num_class, neigh_idx, idx_stride, k, ball_r, B, C, N,
                                 score, ws, ws_bytes, (cudaStream_t)stream);
}

int dsir_topk_rows(const float *score, int B, int N, int k, float *values, int64_t *index, dsir_stream_t stream) {
    if (!score || !values || !index || B <= 0 || N <= 0 || k <= 0 || k > N) return DSIR_ERR_BAD_ARG;
    if (k > 16384 || N >= (1 << 30)) return DSIR_ERR_UNSUPPORTED;
    return launch_topk_rows(score, B, N, k, values, index, (cudaStream_t)stream);
}

int dsir_pose_errors(const float *T_pred, const float *T_gt, int B, float rte_thresh, float rre_thresh, float *out,
                     int32_t *success, dsir_stream_t stream) {
    if (!T_pred || !T_gt || !out || B <= 0) return DSIR_ERR_BAD_ARG;
    return launch_pose_errors(T_pred, T_gt, B, rte_thresh, rre_thresh, out, success, (cudaStream_t)stream);
}

size_t dsir_correspondence_check_workspace_bytes(int64_t total_pos) {
    return total_pos < 0 ? 0 : correspondence_check_workspace_bytes(total_pos);
}

int dsir_correspondence_check(const int32_t *pos_pairs, const int64_t *pos_offsets, int64_t total_pos,
                              const int32_t *pred_pairs, int B, int N, const int64_t *hash_seed, uint8_t *correct, void *ws,
                              size_t ws_bytes, dsir_stream_t stream) {
    if (!pos_offsets || !pred_pairs || !hash_seed || !correct || B <= 0 || N <= 0 || total_pos < 0 || (total_pos > 0 && !pos_pairs))
        return DSIR_ERR_BAD_ARG;
    return launch_correspondence_check(pos_pairs, pos_offsets, total_pos, pred_pairs, B, N, hash_seed, correct, ws, ws_bytes,
                                       (cudaStream_t)stream);
}

size_t dsir_nn_sqdist_workspace_bytes(int B) { return B <= 0 ? 0 : nn_sqdist_workspace_bytes(B); }

int dsir_nn_sqdist_mean(const float *a, const float *b, int B, int N, int M, float *min_d, float *mean, void *ws,
                        size_t ws_bytes, dsir_stream_t stream) {
    if (!a || !b || B <= 0 || N <= 0 || M <= 0 || (!min_d && !mean)) return DSIR_ERR_BAD_ARG;
    return launch_nn_sqdist_mean(a, b, B, N, M, min_d, mean, ws, ws_bytes, (cudaStream_t)stream);
}

}  // extern "C"
