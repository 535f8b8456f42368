// j2k_b200.cu — host side of libj2kb200.so: contexts, plans, launch logic and the C ABI of
// include/j2k_b200.h.  No torch, no CPU fallback: every compute entry point needs a CUDA device.
//
// Data layout in HBM (per device):
//   pixels  : caller frames, interleaved samples (u8 / u16 LE)
//   coeffs  : caller coefficient planes, per frame [tile][component][th][tw] int32 (Mallat layout)
//   scratch : LL ping-pong buffers A (LL of odd levels, <= 1/4 plane) and B (even levels, <= 1/16):
//             the shrinking LL band never lands on unread inputs of another CTA, HL/LH/HH go straight
//             to the coefficient plane, only LL_L is written into the plane.
//   temp    : planar int32/float32 image planes, only on the generic path (Part-2 MCT, planar API,
//             tiles that perform no level).
#include <cuda_runtime.h>

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/j2k_b200.h"
#include "j2k_kernels.cuh"
#include "j2k_pointwise.cuh"
#include "j2k_ring.cuh"
#include "j2k_ht.cuh"
#include "j2k_ht_enc.cuh"

#ifndef J2K_LAUNCH
#define J2K_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define J2K_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

using namespace j2k;

namespace {

// Error text lives in two places: the calling thread (context-free calls: j2k_init, the size / table helpers) and the
// context the failing call ran on.  A cgo caller's goroutine may be moved to another OS thread between the failing call
// and j2k_last_error(); the per-context copy (and j2k_last_error_copy) is what makes the message survive that.
thread_local std::string t_err = "";
thread_local j2k_ctx* t_cur_ctx = nullptr;   // context of the API call running on this thread (CtxGuard)
void publish_error(j2k_ctx* ctx, const char* msg);

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_err = buf;
    if (t_cur_ctx) publish_error(t_cur_ctx, buf);
    return code;
}

struct CtxGuard {  // first statement of every entry point that takes a context
    j2k_ctx* prev;
    explicit CtxGuard(j2k_ctx* c) : prev(t_cur_ctx) { if (c) t_cur_ctx = c; }
    ~CtxGuard() { t_cur_ctx = prev; }
};

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(J2K_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

// ------------------------------------------------------------------ geometry (parity.go, layout.go)

inline int split_len(int n, bool even) { return even ? (n + 1) / 2 : n / 2; }   // wavelet/parity.go:3-8
inline int next_coord(int v) { return (v + 1) >> 1; }                            // wavelet/parity.go:14-16
inline int ceil_div(int a, int b) { return b <= 0 ? 0 : (a >= 0 ? (a + b - 1) / b : a / b); }  // tile_assembler.go:207-215

struct LevelGeom { int w, h, px, py, lw, lh; };
struct Rect { int x, y, w, h; };

// Levels the reference actually performs (ForwardMultilevel*WithParity, dwt53.go:365-394, dwt97.go:388-407:
// stop when the window is <= 1x1) and the residual LL window.
void performed_levels(int w, int h, int x0, int y0, int L, std::vector<LevelGeom>& out, int& cw, int& ch) {
    out.clear();
    cw = w; ch = h;
    int cx = x0, cy = y0;
    for (int l = 0; l < L; l++) {
        if (cw <= 1 && ch <= 1) break;
        if (cw <= 0 || ch <= 0) break;  // empty window (odd origin, 1-sample dimension): every deeper level is a no-op
        LevelGeom g;
        g.w = cw; g.h = ch; g.px = cx & 1; g.py = cy & 1;
        g.lw = split_len(cw, g.px == 0); g.lh = split_len(ch, g.py == 0);
        out.push_back(g);
        cw = g.lw; ch = g.lh; cx = next_coord(cx); cy = next_coord(cy);
    }
}

// bandInfosForResolution over all resolutions, QCD order (encoder.go:2352-2389, t2/geometry.go:53-92)
std::vector<Rect> go_band_rects(int width, int height, int x0, int y0, int L) {
    auto dims = [&](int res, int& rw, int& rh) {
        int level_no = L - res;
        if (level_no < 0) level_no = 0;
        rw = width; rh = height;
        int rx = x0, ry = y0;
        for (int i = 0; i < level_no; i++) {
            rw = split_len(rw, (rx & 1) == 0); rh = split_len(rh, (ry & 1) == 0);
            rx = next_coord(rx); ry = next_coord(ry);
        }
    };
    std::vector<Rect> r;
    int rw, rh;
    dims(0, rw, rh);
    r.push_back({0, 0, rw, rh});
    for (int res = 1; res <= L; res++) {
        int lw, lh;
        dims(res, rw, rh);
        dims(res - 1, lw, lh);
        int hw = rw - lw, hh = rh - lh;
        r.push_back({lw, 0, hw, lh});
        r.push_back({0, lh, lw, hh});
        r.push_back({lw, lh, hw, hh});
    }
    return r;
}

inline bool rect_empty(const Rect& r) { return r.w <= 0 || r.h <= 0; }
inline bool rect_same(const Rect& a, const Rect& b) {
    if (rect_empty(a) && rect_empty(b)) return true;
    return a.x == b.x && a.y == b.y && a.w == b.w && a.h == b.h;
}

// True when the literal band rectangles coincide with the windows the DWT levels write, so the
// quantizer can be fused into the level kernels (always, except odd origins with 1-sample windows).
bool geometry_regular(const std::vector<LevelGeom>& lv, int cw, int ch, const std::vector<Rect>& rects, int L) {
    std::vector<Rect> mine(3 * L + 1, Rect{0, 0, 0, 0});
    mine[0] = {0, 0, cw, ch};
    for (size_t k = 0; k < lv.size(); k++) {  // level k+1 <-> resolution L-k
        const LevelGeom& g = lv[k];
        int res = L - (int)k;
        int idx = 3 * (res - 1) + 1;
        mine[idx] = {g.lw, 0, g.w - g.lw, g.lh};
        mine[idx + 1] = {0, g.lh, g.lw, g.h - g.lh};
        mine[idx + 2] = {g.lw, g.lh, g.w - g.lw, g.h - g.lh};
    }
    for (int i = 0; i < 3 * L + 1; i++)
        if (!rect_same(mine[i], rects[i])) return false;
    return true;
}

// ------------------------------------------------------------------ device resources

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return fail(J2K_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

enum BufId { B_NONE = 0, B_PIX, B_COEF, B_PLANES, B_TEMP, B_SA, B_SB, B_CTEMP, B_COUNT };

struct LevelLaunch {
    LevelArgs a;
    int WT, NP, NC, KIND, MCT;  // KIND: InKind of the interleaved side
    int x_buf, ll_buf, band_buf, planes_buf;
    int x_elem_bytes;           // element size of x_base for alignment checks
    int level;                  // 1 = finest
    bool fast;                  // eligible for the fast-path kernel (FwdFast)
    FastQ4 fq;
    int cls;                    // tile class the launch belongs to
    bool nc3_first;             // 3-component image-side launch (items are pixel items)
    size_t o_x, o_ll, o_plane;  // offset-table positions (alignment checks of the ring path)
    bool in_ring = false;       // handled by the persistent launch (levels 1..RingPlan::cut)
};

// Persistent single-launch form of a whole plan direction (j2k_ring.cuh); built when every level qualifies.
struct RingPlan {
    bool ok = false;
    RingArgs args;
    int WT = 0, NP1 = 0, NC1 = 0, IN1 = 0, MCT1 = 0, SG1 = 0;
    int UA = 0;  // general-alignment kernel variant (rows / bands that are not 16-byte aligned, partial vectors at the right edge)
    int X3 = 0;  // inverse, ICT + 9/7 8-bit RGB: the three-producer level-1 kernel (inv3w_kernel), jobs claimed three at a time
    int x_buf[J2K_RING_MAXSEG], ll_buf[J2K_RING_MAXSEG], band_buf[J2K_RING_MAXSEG], planes_buf[J2K_RING_MAXSEG];
    int level[J2K_RING_MAXSEG];
    int cut = 0;  // levels 1..cut run in the persistent launch, deeper ones (geometry it does not take) on the per-level kernels
    unsigned grid = 0;
};

enum PwKind { PW_PREP, PW_FINALIZE, PW_QUANT_RECTS, PW_SHIFT, PW_COPY, PW_DEQUANT_RECTS };
struct PwLaunch {
    int kind;
    int src_buf, dst_buf;
    const long long *src_off, *dst_off;
    int src_stride, dst_stride, n_items, w, h, cvt, all_round, shift;
    RectTable rt;
};

struct Plan;

// Pinned staging of one device for caller-owned PAGEABLE buffers (SURVEY 8b "Ownership": a Go []byte from
// PixelData.GetFrame is pageable; the synchronous calls move it through C-owned pinned memory).  A ring of chunks per
// direction: host threads memcpy a chunk while the copy engine moves the previous ones, so the upload of sub-batch b+1,
// the kernels of b and the download of b-1 overlap exactly as they do for pinned callers.
struct Stager {
    size_t CHUNK = 4u << 20;   // bytes per staging chunk (env J2K_STAGE_CHUNK_KB: tests use small chunks)
    static constexpr int NSLOT = 8;
    unsigned char* up[NSLOT] = {};
    unsigned char* down[NSLOT] = {};
    cudaEvent_t up_ev[NSLOT] = {}, down_ev[NSLOT] = {};
    bool up_busy[NSLOT] = {};
    int up_next = 0;
    bool ready = false;
    struct Pending { unsigned char* dst; const unsigned char* src; size_t bytes; };  // device -> pageable host, not yet issued
    std::vector<Pending> pending;
};

struct DeviceCtx {
    int dev = 0;
    bool failed = false;   // a CUDA call failed and the device did not come back: left out of the round-robin (run_host_batch)
    Stager stg;
    cudaStream_t s_main = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_t[4] = {nullptr, nullptr, nullptr, nullptr};      // timing
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    bool ev_k_used[2] = {false, false}, ev_out_used[2] = {false, false};
    DevBuf in[2], out[2], planes[2], api[8];
    DevBuf blk[2], nbp[2];  // code-block interface: block-major planes and per-block numbps of a sub-batch
    DevBuf gsh[2], gmk[2];  // general-scaling ROI of a sub-batch: per-block shifts, per-sample mask
    DevBuf ht_bytes[2], ht_desc[2], ht_status[2];  // HTJ2K block decoding of a sub-batch: cleanup segments, records, result codes
    DevBuf ht_scratch;                             // (inf, u_q) per quad between the two HT kernels
    cudaEvent_t ev_ht = nullptr; bool ev_ht_used = false;
    // HTJ2K block encoding: per-block scratch slots, block info, stream offsets, per-block Kmax; per sub-batch slot the compact
    // stream + records on the device and a pinned landing area for (total, records)
    DevBuf he_slots, he_info, he_off, he_kmax, he_bytes[2], he_recs[2];
    std::vector<unsigned char> he_kmax_host;
    void* he_pin[2] = {nullptr, nullptr}; size_t he_pin_cap[2] = {0, 0};
    std::map<std::string, std::unique_ptr<Plan>> plans;
    long long use_clock = 0;   // LRU stamps of `plans`
    std::map<std::string, std::unique_ptr<struct BlockTable>> block_tables;
};

struct Plan {
    bool fwd = true;
    int nframes = 0;
    // shared
    int C = 1, W = 0, H = 0, L = 0, bps = 1;
    bool reversible = true;
    bool generic = false;        // prep / finalize path
    bool direct = false;         // wavelet package API: planar working-type in, raw bands out
    RawFmt raw{};
    MctProgram prog{};
    int pix_kind = IN_U8;
    long long frame_samples = 0; // caller frame stride in samples
    long long coeffs_per_frame = 0;
    bool want_planes = false;
    DevBuf tables, temp, sa, sb, ctemp, ctl;
    RingPlan ring;
    std::vector<LevelLaunch> levels;   // in execution order
    std::vector<PwLaunch> pre, post;   // pointwise launches before / after the level launches
    int launches_per_run = 0;
    long long last_use = 0;
    ~Plan() { tables.release(); temp.release(); sa.release(); sb.release(); ctemp.release(); ctl.release(); }
};

}  // namespace

struct j2k_ctx {
    std::vector<DeviceCtx> devs;
    std::mutex mu;
    std::mutex err_mu;
    std::string last_err;       // message of the most recent failing call on this context (any thread)
    long long err_seq = 0;      // number of failures so far
    // host threads that fill / drain the pinned staging chunks of pageable callers (started on first use)
    std::vector<std::thread> workers;
    std::mutex wq_mu;
    std::condition_variable wq_cv, wq_done;
    std::deque<std::function<void()>> wq;
    int wq_inflight = 0;
    bool wq_stop = false;
    std::atomic<long long> launches{0};
    j2k_timing last{};
    std::vector<void*> pinned;
    long long next_ticket = 1;
    struct TicketEv { int di; cudaEvent_t ev; };
    std::map<long long, std::vector<TicketEv>> tickets;  // per device: an event behind the job's last D2H copy
    // per-launch profiling (j2k_set_profiling)
    bool profiling = false;
    int prof_dev = 0;
    std::vector<cudaEvent_t> prof_ev;   // 2 per launch
    std::vector<int> prof_level;
    cudaStream_t prof_stream = nullptr;
};

namespace {

void publish_error(j2k_ctx* ctx, const char* msg) {
    std::lock_guard<std::mutex> lk(ctx->err_mu);
    ctx->last_err = msg;
    ctx->err_seq++;
}

// ------------------------------------------------------------------ level kernel dispatch

#define J2K_GRID(a) (unsigned)((((long long)(a).n_items * (a).nchunks * (a).nstrips) + 3) / 4)

#define FWD_CASE(wt, np, nc, in, mct)                                                                   \
    if (l.WT == wt && l.NP == np && l.NC == nc && l.KIND == in && l.MCT == mct) {                       \
        J2K_LAUNCH((fwd_level_kernel<wt, np, nc, in, mct>), J2K_GRID(a), 128, st, a);                   \
        return cudaGetLastError();                                                                      \
    }
#define INV_CASE(wt, np, nc, out, mct)                                                                  \
    if (l.WT == wt && l.NP == np && l.NC == nc && l.KIND == out && l.MCT == mct) {                      \
        J2K_LAUNCH((inv_level_kernel<wt, np, nc, out, mct>), J2K_GRID(a), 128, st, a);                  \
        return cudaGetLastError();                                                                      \
    }

cudaError_t launch_fwd_level(const LevelLaunch& l, const LevelArgs& a, cudaStream_t st) {
    FWD_CASE(53, 4, 1, IN_U8, MCTK_NONE) FWD_CASE(53, 4, 1, IN_U16, MCTK_NONE)
    FWD_CASE(53, 1, 1, IN_U8, MCTK_NONE) FWD_CASE(53, 1, 1, IN_U16, MCTK_NONE)
    FWD_CASE(53, 1, 1, IN_I32, MCTK_NONE) FWD_CASE(53, 2, 1, IN_I32, MCTK_NONE)
    FWD_CASE(53, 2, 3, IN_U8, MCTK_RCT) FWD_CASE(53, 2, 3, IN_U16, MCTK_RCT)
    FWD_CASE(97, 4, 1, IN_U8, MCTK_NONE) FWD_CASE(97, 4, 1, IN_U16, MCTK_NONE)
    FWD_CASE(97, 1, 1, IN_U8, MCTK_NONE) FWD_CASE(97, 1, 1, IN_U16, MCTK_NONE)
    FWD_CASE(97, 1, 1, IN_I32, MCTK_NONE) FWD_CASE(97, 2, 1, IN_I32, MCTK_NONE)
    FWD_CASE(97, 1, 1, IN_F32, MCTK_NONE) FWD_CASE(97, 2, 1, IN_F32, MCTK_NONE)
    FWD_CASE(97, 2, 3, IN_U8, MCTK_ICT) FWD_CASE(97, 2, 3, IN_U16, MCTK_ICT)
    return cudaErrorInvalidDeviceFunction;
}

#define FAST_CASE(wt, np, nc, in, mct)                                                                  \
    if (l.WT == wt && l.NP == np && l.NC == nc && l.KIND == in && l.MCT == mct) {                       \
        J2K_LAUNCH((fwd_fast_kernel<wt, np, nc, in, mct>), J2K_GRID(a), 128, st, a, l.fq);              \
        return cudaGetLastError();                                                                      \
    }

bool has_fwd_fast(const LevelLaunch& l) {
    if (l.NC == 3) return l.NP == 2 && (l.KIND == IN_U8 || l.KIND == IN_U16);
    if (l.KIND == IN_U8 || l.KIND == IN_U16) return l.NP == 4;
    return l.NP == 2;
}

cudaError_t launch_fwd_fast(const LevelLaunch& l, const LevelArgs& a, cudaStream_t st) {
    FAST_CASE(53, 4, 1, IN_U8, MCTK_NONE) FAST_CASE(53, 4, 1, IN_U16, MCTK_NONE) FAST_CASE(53, 2, 1, IN_I32, MCTK_NONE)
    FAST_CASE(53, 2, 3, IN_U8, MCTK_RCT) FAST_CASE(53, 2, 3, IN_U16, MCTK_RCT)
    FAST_CASE(97, 4, 1, IN_U8, MCTK_NONE) FAST_CASE(97, 4, 1, IN_U16, MCTK_NONE) FAST_CASE(97, 2, 1, IN_I32, MCTK_NONE)
    FAST_CASE(97, 2, 1, IN_F32, MCTK_NONE)
    FAST_CASE(97, 2, 3, IN_U8, MCTK_ICT) FAST_CASE(97, 2, 3, IN_U16, MCTK_ICT)
    return cudaErrorInvalidDeviceFunction;
}

cudaError_t launch_inv_level(const LevelLaunch& l, const LevelArgs& a, cudaStream_t st) {
    INV_CASE(53, 4, 1, IN_U8, MCTK_NONE) INV_CASE(53, 4, 1, IN_U16, MCTK_NONE)
    INV_CASE(53, 1, 1, IN_U8, MCTK_NONE) INV_CASE(53, 1, 1, IN_U16, MCTK_NONE)
    INV_CASE(53, 1, 1, IN_I32, MCTK_NONE) INV_CASE(53, 2, 1, IN_I32, MCTK_NONE)
    INV_CASE(53, 2, 3, IN_U8, MCTK_RCT) INV_CASE(53, 2, 3, IN_U16, MCTK_RCT)
    INV_CASE(97, 4, 1, IN_U8, MCTK_NONE) INV_CASE(97, 4, 1, IN_U16, MCTK_NONE)
    INV_CASE(97, 1, 1, IN_U8, MCTK_NONE) INV_CASE(97, 1, 1, IN_U16, MCTK_NONE)
    INV_CASE(97, 1, 1, IN_F32, MCTK_NONE) INV_CASE(97, 2, 1, IN_F32, MCTK_NONE)
    INV_CASE(97, 2, 3, IN_U8, MCTK_ICT) INV_CASE(97, 2, 3, IN_U16, MCTK_ICT)
    return cudaErrorInvalidDeviceFunction;
}

// ------------------------------------------------------------------ plan building

struct TileGeom { int x0, y0, tw, th, ox, oy; long long coeff_off; };

struct Spec {  // superset of the public parameter blocks
    bool fwd;
    int W, H, C, bit_depth, is_signed, L, reversible, htj2k, mct_mode;
    std::vector<TileGeom> tiles;
    int n_steps; double steps[J2K_MAX_BANDS];
    bool fuse_shift;        // fwd <<6 / inv /2
    bool planar_in;         // j2k_forward_planar
    bool direct;            // wavelet API
    bool want_planes;
    const double* mct_matrix; int mct_has_offsets; const int32_t* mct_offsets;
    int n_bindings; const j2k_mct_binding* bindings;
};

RawFmt make_raw(const Spec& s) {
    RawFmt r{};
    r.pix_stride = s.C;
    if (s.is_signed) {
        if (s.bit_depth <= 8) { r.sign_thresh = 128; r.sign_sub = 256; }               // encoder.go:362-364
        else { r.sign_thresh = 1 << (s.bit_depth - 1); r.sign_sub = 1 << s.bit_depth; } // encoder.go:374-376
        r.dc = 0;
        r.clamp_lo = -(1 << (s.bit_depth - 1)); r.clamp_hi = (1 << (s.bit_depth - 1)) - 1;
        r.wrap_add = 1 << s.bit_depth;
    } else {
        r.sign_thresh = 0x7fffffff; r.sign_sub = 0;
        r.dc = 1 << (s.bit_depth - 1);
        r.clamp_lo = 0; r.clamp_hi = (1 << s.bit_depth) - 1;
        r.wrap_add = 0;
    }
    return r;
}

void fill_op_from_binding(const j2k_mct_binding& b, int C, bool fwd, MctOp& op) {
    memset(&op, 0, sizeof op);
    int n = b.n_components;
    if (n == 0 && fwd) { n = C; for (int i = 0; i < n; i++) op.ids[i] = i; }  // encoder.go:562-568
    else for (int i = 0; i < n; i++) op.ids[i] = b.component_ids[i];
    op.n = n;
    op.kind = b.element_type == 0 ? 0 : 1;
    op.has_matrix = fwd ? 1 : b.has_matrix;
    for (int r = 0; r < n; r++)
        for (int k = 0; k < n; k++) {
            double m = b.has_matrix ? b.matrix[r * n + k] : (r == k ? 1.0 : 0.0);  // encoder.go:596-608
            op.mf[r * n + k] = m;
            if (fwd && op.kind == 1) op.mi[r * n + k] = (int)(m * (double)(1 << 13));
            else op.mi[r * n + k] = (int)m;
        }
    op.has_off = b.has_offsets;
    for (int i = 0; i < n; i++) op.off[i] = b.offsets[i];
}

int make_prog(const Spec& s, MctProgram& prog) {
    memset(&prog, 0, sizeof prog);
    switch (s.mct_mode) {
    case J2K_MCT_NONE: prog.kind = 0; return 0;
    case J2K_MCT_RCT:
        if (s.C != 3 || !s.reversible) return fail(J2K_ERR_INVALID_ARG, "RCT needs 3 components and the reversible transform");
        prog.kind = 1; return 0;
    case J2K_MCT_ICT:
        if (s.C != 3 || s.reversible) return fail(J2K_ERR_INVALID_ARG, "ICT needs 3 components and the irreversible transform");
        prog.kind = 2; return 0;
    case J2K_MCT_CUSTOM_INT: case J2K_MCT_CUSTOM_Q13: case J2K_MCT_CUSTOM_FLOAT: {
        if ((s.mct_mode == J2K_MCT_CUSTOM_FLOAT) == s.fwd) return fail(J2K_ERR_INVALID_ARG, "custom MCT mode %d is not valid in this direction", s.mct_mode);
        prog.kind = 3; prog.n_ops = 1;
        MctOp& op = prog.ops[0];
        op.n = s.C;
        for (int i = 0; i < s.C; i++) op.ids[i] = i;
        op.kind = s.mct_mode == J2K_MCT_CUSTOM_INT ? 0 : 1;
        op.has_matrix = 1;
        for (int i = 0; i < s.C * s.C; i++) {
            op.mf[i] = s.mct_matrix[i];
            op.mi[i] = s.mct_mode == J2K_MCT_CUSTOM_Q13 ? (int)(s.mct_matrix[i] * (double)(1 << 13)) : (int)s.mct_matrix[i];
        }
        op.has_off = s.mct_has_offsets;
        for (int i = 0; i < s.C; i++) op.off[i] = s.mct_offsets[i];
        return 0;
    }
    case J2K_MCT_BINDINGS:
        if (s.n_bindings < 0 || s.n_bindings > J2K_MAX_BINDINGS) return fail(J2K_ERR_INVALID_ARG, "n_bindings out of range");
        prog.kind = 3; prog.n_ops = 0;
        for (int i = 0; i < s.n_bindings; i++) {
            const j2k_mct_binding& b = s.bindings[i];
            if (b.n_components < 0 || b.n_components > s.C) return fail(J2K_ERR_INVALID_ARG, "binding %d: bad component count", i);
            for (int k = 0; k < b.n_components; k++)
                if (b.component_ids[k] < 0 || b.component_ids[k] >= s.C) return fail(J2K_ERR_INVALID_ARG, "binding %d: bad component id", i);
            if (!s.fwd && b.n_components == 0) continue;  // decoder.go:633-635
            fill_op_from_binding(b, s.C, s.fwd, prog.ops[prog.n_ops++]);
        }
        return 0;
    }
    return fail(J2K_ERR_INVALID_ARG, "unknown mct_mode %d", s.mct_mode);
}

inline bool all_mult(const std::vector<long long>& v, size_t b, size_t e, int m) {
    for (size_t i = b; i < e; i++) if (v[i] % m) return false;
    return true;
}

struct TableBuilder {
    std::vector<long long> host;
    size_t add(const std::vector<long long>& v) { size_t o = host.size(); host.insert(host.end(), v.begin(), v.end()); return o; }
};

struct ClassTmp {
    std::vector<int> tiles;
    std::vector<LevelGeom> lv;
    int cw, ch, tw, th, ox, oy;
    bool regular;
    std::vector<Rect> rects;
};

void choose_chunks(LevelArgs& a, int valid_pairs) {
    a.nstrips = (a.Kx + valid_pairs - 1) / valid_pairs;
    if (a.nstrips < 1) a.nstrips = 1;
    long long base = (long long)a.n_items * a.nstrips;
    const long long target = 148LL * 32;  // >= 2 waves of 16 warps per SM
    long long want = (target + base - 1) / base;
    int max_chunks = (a.Ky + 15) / 16;
    if (max_chunks < 1) max_chunks = 1;
    if (want > max_chunks) want = max_chunks;
    if (want < 1) want = 1;
    a.chunk_pairs = (int)((a.Ky + want - 1) / want);
    if (a.chunk_pairs < 1) a.chunk_pairs = 1;
    a.nchunks = (a.Ky + a.chunk_pairs - 1) / a.chunk_pairs;
    if (a.nchunks < 1) a.nchunks = 1;
}

inline int valid_pairs_of(int wt, int np) {
    int halo = wt == 97 ? 2 : 1;
    int hln = (halo + np - 1) / np;
    return (32 - 2 * hln) * np;
}

void set_geom(LevelArgs& a, const LevelGeom& g) {
    a.w = g.w; a.h = g.h; a.px = g.px; a.py = g.py; a.lw = g.lw; a.lh = g.lh;
    a.Kx = (g.w + g.px + 1) / 2; a.Ky = (g.h + g.py + 1) / 2;
    a.hskip = g.w <= 1; a.vskip = g.h <= 1;
}

// The quantizer computes the IEEE quotient c / step' from the correctly rounded reciprocal (Markstein: q0 = c * r,
// e = fma(-q0, step', c), q = fma(e, r, q0); tests/test_div_markstein.py).  That is exact when step' is a normal float32
// whose significand is not all ones (the divisor class the published theorem excludes) and whose exponent leaves room
// for q0 and e next to |c| in {0} U [2^-40, 2^30] (coefficients of <= 16-bit samples).  Runtime steps have 11 fraction bits
// (quantization.go:130-154), so every step the reference produces qualifies; any other step takes the IEEE division.
bool markstein_safe(float step_eff) {
    uint32_t u;
    memcpy(&u, &step_eff, 4);
    const int ex = (int)((u >> 23) & 0xFF) - 127;
    return (u >> 31) == 0 && ex >= -40 && ex <= 40 && (u & 0x7FFFFFu) != 0x7FFFFFu;
}

// Quantizer / dequantizer mode of band `idx` (QCD order).
void band_mode_fwd(const Spec& s, int idx, BandIO& b) {
    b.shift = 0; b.step = 1.f; b.rcp = 1.f; b.scale = 1.f;
    if (s.direct) { b.mode = Q_RAW; return; }
    if (s.reversible) {
        if (s.fuse_shift && !s.htj2k) { b.mode = Q_SHIFT; b.shift = 6; } else b.mode = Q_RAW;
        return;
    }
    if (s.n_steps == 0) { b.mode = Q_ROUND; return; }     // encoder.go:2266-2273
    if (idx >= s.n_steps) { b.mode = Q_ZERO; return; }    // encoder.go:2283,2294
    if (s.steps[idx] <= 0) { b.mode = Q_ROUND; return; }  // encoder.go:2320-2321
    b.mode = Q_QUANT;
    b.step = (float)s.steps[idx];                          // float32(stepSize), encoder.go:2323
    b.scale = s.htj2k ? 1.f : 64.f;                        // encoder.go:2312-2315
    b.rcp = markstein_safe(b.step) && markstein_safe(b.step / b.scale) ? 1.0f / b.step : 0.f;  // 0: div_by_step divides
}
void band_mode_inv(const Spec& s, int idx, BandIO& b) {
    b.shift = 0; b.step = 1.f; b.rcp = 1.f; b.scale = 1.f;
    if (s.direct) { b.mode = DQ_RAW; return; }
    if (s.reversible) { b.mode = (s.fuse_shift && !s.htj2k) ? DQ_HALVE : DQ_RAW; return; }
    if (s.n_steps == 0 || idx >= s.n_steps || s.steps[idx] <= 0) { b.mode = DQ_CVT; return; }
    b.mode = DQ_SCALE;
    b.scale = (float)(s.htj2k ? s.steps[idx] : 0.5 * s.steps[idx]);  // t2/tile_decoder.go:974-977,983
}

void fill_rect_table(const Spec& s, const std::vector<Rect>& rects, RectTable& rt) {
    memset(&rt, 0, sizeof rt);
    rt.n = (int)rects.size();
    for (int i = 0; i < rt.n; i++) {
        rt.x[i] = rects[i].x; rt.y[i] = rects[i].y; rt.w[i] = rects[i].w; rt.h[i] = rects[i].h;
        BandIO b{};
        if (s.fwd) band_mode_fwd(s, i, b); else band_mode_inv(s, i, b);
        rt.mode[i] = b.mode; rt.step[i] = b.step; rt.rcp[i] = b.rcp; rt.scale[i] = b.scale;
    }
}


// ------------------------------------------------------------------ persistent single-launch plan (j2k_ring.cuh)

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// SMs of the current device (the emulator has one)
int device_sm_count() {
#ifdef J2K_EMU
    return 1;
#else
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) return 148;
    return sms;
#endif
}

// Job decomposition of one ring segment: even column strips, row chunks sized for ~2 jobs per resident warp.
// target_halves: job target in halves of the default (3 = 1.5 x); chunk_cap: upper bound of the chunk height (0 = the default)
void ring_chunks(RingSeg& g, int NP, int n_cols_total_hint, int level, int halo_lanes = 2, int target_div = 1, int target_halves = 2,
                 int chunk_cap = 0) {
    int VP = (32 - halo_lanes) * NP;  // 32 lanes minus one halo lane per side (9/7 needs 2 pairs, 5/3 one: both fit NP >= 2); halo-free: 32
    // strips are a multiple of 8 pairs wide: every band-row piece a warp stores (and every pixel-row piece of the inverse)
    // then covers whole 32-byte sectors, no partial-sector writes at the strip seams (+3 % on C2)
    int align = 8 > NP ? 8 : NP;
    {   // experiment knobs: strip width limit and alignment in pairs
        int m = env_int("J2K_RING_STRIP_MAX", 0), al = env_int("J2K_RING_STRIP_ALIGN", 0);
        if (m > 0 && m < VP) VP = m / NP * NP;
        if (al >= NP) align = al / NP * NP;
    }
    g.nstrips = (g.Kx + VP - 1) / VP;
    if (g.nstrips < 1) g.nstrips = 1;
    int sp = (g.Kx + g.nstrips - 1) / g.nstrips;
    g.strip_pairs = (sp + align - 1) / align * align;
    if (g.strip_pairs > VP) g.strip_pairs = VP;
    g.nstrips = (g.Kx + g.strip_pairs - 1) / g.strip_pairs;
    // chunks as tall as the job target allows, up to 128 row pairs: the 2 (5/3) or 4 (9/7) warm-up pairs a chunk recomputes are
    // then 1.5-3 % of its rows (6 % at 64); the difference only shows once the board is power-capped
    int max_chunk = level <= 1 ? env_int("J2K_RING_CHUNK", 128) : env_int("J2K_RING_CHUNK_DEEP", env_int("J2K_RING_CHUNK", 128));
    if (chunk_cap > 0 && chunk_cap < max_chunk) max_chunk = chunk_cap;
    // (target_div: a job of the three-producer inverse occupies a whole CTA, and every job boundary drains its exchange
    // pipeline: four times fewer, taller jobs: +11 % on C3, +2.5 % on C5)
    const long long target = (long long)env_int("J2K_RING_TARGET_JOBS", 148 * 16 * 2) * target_halves / 2 / target_div;
    long long cols = (long long)n_cols_total_hint * g.nstrips;
    long long want = (target + cols - 1) / cols;
    if (want < 1) want = 1;
    int cp = (int)((g.Ky + want - 1) / want);
    const int min_chunk = env_int("J2K_RING_CHUNK_MIN", 8);
    if (cp < min_chunk) cp = min_chunk;
    if (cp > max_chunk) cp = max_chunk;
    if (cp > g.Ky) cp = g.Ky;
    if (cp < 1) cp = 1;
    g.chunk_pairs = cp;
    g.nchunks = (g.Ky + cp - 1) / cp;
}

// Level-1 chunk height by estimated makespan: the job count should fill whole rounds of the `slots` warps that run such jobs
// at once; cost of n chunks of cp row pairs = rounds x (cp + warm-up pairs + start-up bubble).
void ring_chunks_makespan(RingSeg& g, long long cols, long long slots, int warm_pairs, int max_chunk) {
    const int min_chunk = env_int("J2K_RING_CHUNK_MIN", 8), bubble = env_int("J2K_RING_BUBBLE", 8);
    long long best = -1;
    for (int n = 1; n <= g.Ky; n++) {
        const int cp = (g.Ky + n - 1) / n;
        if (cp > max_chunk) continue;
        if (cp < min_chunk && best >= 0) break;
        const int nch = (g.Ky + cp - 1) / cp;
        const long long nj = cols * nch;
        const long long cost = ((nj + slots - 1) / slots) * (cp + warm_pairs + bubble);
        if (best < 0 || cost < best) { best = cost; g.chunk_pairs = cp; g.nchunks = nch; }
    }
}

bool ring_variant_supported(int WT, int NP, int NC, int IN, int MCT, int SG) {
    if (NC == 3 && NP == 4) return WT == 97 && SG == 0 && (IN == IN_U8 || IN == IN_U16) && MCT == MCTK_ICT;  // component-split first level
    if (NC == 3) return NP == 2 && SG == 0 && (IN == IN_U8 || IN == IN_U16) && MCT == (WT == 53 ? MCTK_RCT : MCTK_ICT);
    if (NP != 4 || MCT != MCTK_NONE) return false;
    if (IN == IN_U8 || IN == IN_U16) return true;
    if (SG) return false;
    return IN == IN_I32 || (IN == IN_F32 && WT == 97);
}

// Job order of the persistent launch.  The plain list is level-major: every item finishes level k before any job of
// level k+1 is claimed, so an intermediate LL plane has long left the 126 MB L2 when its consumer reads it.  When the
// segments form one dependency chain (one tile class) the list is cut into slices instead: items are taken in groups of
// ~J2K_RING_GROUP_KS Ki samples, and the slice (level k, group g) is placed at tick g + k * lag, i.e. `lag` groups of
// level-1 work behind the slice that produces its input — far enough that the producers have retired when it is
// claimed (no warp parks on a dependency), close enough that the LL plane is still resident in L2.  A slice's producer
// always precedes it in the list, so the in-order claiming argument (no co-residency requirement) is unchanged.
// Allocates and zeroes the control block; the slice table lives behind the counters.
int ring_schedule(RingPlan& R, Plan& P, int n_ctl) {
    RingArgs& A = R.args;
    std::vector<int> begin, info;
    const int lag = env_int("J2K_RING_LAG", J2K_RING_DEFAULT_LAG);
    bool chain = A.nseg >= 2 && lag > 0 && !R.X3;  // (the three-producer inverse numbers its jobs in triples: level-major order only)
    int base = 0x7fffffff;
    for (int k = 0; k < A.nseg; k++) base = A.seg[k].n_items < base ? A.seg[k].n_items : base;
    for (int k = 0; k < A.nseg && chain; k++)
        if (base < 1 || A.seg[k].n_items % base || A.seg[k].dep_seg != k - 1) chain = false;
    if (chain) {
        long long big = 1;  // samples one base item (a frame / tile with all its components) holds at the largest level
        for (int k = 0; k < A.nseg; k++) {
            long long sz = (long long)A.seg[k].w * A.seg[k].h * (A.seg[k].n_items / base) * (A.seg[k].first && R.NC1 == 3 && R.NP1 == 2 ? 3 : 1);
            big = sz > big ? sz : big;
        }
        const long long target = (long long)env_int("J2K_RING_GROUP_KS", J2K_RING_DEFAULT_GROUP_KS) * 1024;
        long long G = target / big;
        if (G < 1) G = 1;
        const long long max_groups = J2K_RING_MAXSLICE / A.nseg;
        if ((base + G - 1) / G > max_groups) G = (base + max_groups - 1) / max_groups;
        const int NG = (int)((base + G - 1) / G);
        if (NG > lag) {  // fewer groups than the lag: nothing to pipeline
            long long jobs = 0;
            for (int t = 0; t < NG + (A.nseg - 1) * lag; t++)
                for (int k = A.nseg - 1; k >= 0; k--) {
                    const int g = t - k * lag;
                    if (g < 0 || g >= NG) continue;
                    const long long ratio = A.seg[k].n_items / base;
                    const long long i0 = g * G * ratio, i1 = ((g + 1) * G < base ? (g + 1) * G : base) * ratio;
                    begin.push_back((int)jobs);
                    info.push_back(k); info.push_back((int)i0);
                    jobs += (i1 - i0) * A.seg[k].nchunks * A.seg[k].nstrips;
                }
            if (jobs != A.total_jobs) return fail(J2K_ERR_INVALID_ARG, "ring schedule: %lld jobs in slices, %d in segments", jobs, A.total_jobs);
        }
    }
    // L2 policies: whatever is read once or written for the host leaves L2 first; the intermediate LL planes stay
    {
        static const unsigned long long pol[3] = {J2K_L2_EVICT_NORMAL, J2K_L2_EVICT_FIRST, J2K_L2_EVICT_LAST};
        const int on = !begin.empty();
        const int p_in = env_int("J2K_RING_POL_IN", on ? 1 : 0) % 3, p_llr = env_int("J2K_RING_POL_LLR", 0) % 3;
        const int p_llw = env_int("J2K_RING_POL_LLW", on ? 2 : 0) % 3, p_band = env_int("J2K_RING_POL_BAND", on ? 1 : 0) % 3;
        for (int k = 0; k < A.nseg; k++) {
            RingSeg& g = A.seg[k];
            if (P.fwd) {
                g.pol_load = pol[g.first ? p_in : p_llr];
                g.pol_ll = pol[g.has_waiters ? p_llw : p_band];
                g.pol_band = pol[p_band];
            } else {
                g.pol_load = pol[p_in];
                g.pol_ll = pol[p_llr];
                g.pol_band = pol[g.has_waiters ? p_llw : p_band];
            }
        }
    }
    const size_t ctl_ints = ((size_t)n_ctl + 3) / 4 * 4, nb = (begin.size() + 1) / 2 * 2;
    int rc = P.ctl.ensure((ctl_ints + nb + info.size()) * sizeof(unsigned));
    if (rc) return rc;
    CK(cudaMemset(P.ctl.p, 0, (size_t)n_ctl * sizeof(unsigned)));
    A.ctl = (unsigned*)P.ctl.p;
    A.nslice = (int)begin.size();
    A.slice_begin = nullptr; A.slice_info = nullptr;
    if (A.nslice) {
        int* tb = (int*)P.ctl.p + ctl_ints;
        CK(cudaMemcpy(tb, begin.data(), begin.size() * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(tb + nb, info.data(), info.size() * sizeof(int), cudaMemcpyHostToDevice));
        A.slice_begin = tb;
        A.slice_info = (const int2*)(tb + nb);
    }
    return 0;
}

// Converts the per-level launch list into ONE persistent launch when every level qualifies; otherwise P.ring.ok stays false
// and run_plan uses the per-level kernels.
int build_ring_fwd_impl(const Spec& s, Plan& P, const std::vector<long long>& tab, bool split3, int cut, bool allow_x3);
int ring_schedule(RingPlan& R, Plan& P, int n_ctl);

// The persistent launch takes levels 1..cut, the largest cut whose levels all have the geometry it needs (window widths
// that are multiples of 8, 16-byte aligned rows); deeper levels, which hold 1/4^cut of the samples, stay on the per-level
// kernels behind it in stream order.  3-component 9/7 frames first try the component-split first level (NP = 4, one
// component per job) and fall back to the three-components-per-job variant (NP = 2).
int build_ring_fwd(const Spec& s, Plan& P, const std::vector<long long>& tab) {
    int maxlevel = 0;
    for (auto& l : P.levels) maxlevel = l.level > maxlevel ? l.level : maxlevel;
    // split 2: component-split level 1 on fwd3w_kernel (big launches only), 1: as component jobs of fwd_ring_kernel, 0: NP = 2
    for (int split = (!s.reversible && env_int("J2K_RING_SPLIT3", 1)) ? (env_int("J2K_FWD3W", J2K_FWD3W_DEFAULT) ? 2 : 1) : 0; split >= 0; split--)
        for (int cut = maxlevel; cut >= 1; cut--) {
            int rc = build_ring_fwd_impl(s, P, tab, split != 0, cut, split == 2);
            if (rc) return rc;
            if (P.ring.ok) {
                P.ring.cut = cut;
                for (auto& l : P.levels) l.in_ring = l.level <= cut;
                return 0;
            }
        }
    return 0;
}

int build_ring_fwd_impl(const Spec& s, Plan& P, const std::vector<long long>& tab, bool split3, int cut, bool allow_x3) {
    RingPlan& R = P.ring;
    R.ok = false;
    if (env_int("J2K_RING_DISABLE", 0)) return 0;
    if (P.levels.empty()) return 0;
    memset(&R.args, 0, sizeof R.args);
    const int WT = s.reversible ? 53 : 97;
    // order: level-major, then class (the launch list is class-major)
    std::vector<int> order;
    for (int k = 1; k <= cut; k++)
        for (size_t i = 0; i < P.levels.size(); i++)
            if (P.levels[i].level == k) order.push_back((int)i);
    if (order.empty() || order.size() > J2K_RING_MAXSEG) return 0;
    std::vector<int> seg_of(P.levels.size(), -1);
    bool have_first = false;
    bool ua = false;
    int n_ctl = 2, jobs = 0;
    R.X3 = 0;
    // component-split level 1 as one converting producer warp + three single-component consumers per CTA (fwd3w_kernel)
    const bool use_x3 = split3 && allow_x3;
    for (size_t si = 0; si < order.size(); si++) {
        const LevelLaunch& l = P.levels[order[si]];
        const LevelArgs& a = l.a;
        RingSeg& g = R.args.seg[si];
        const bool first = l.level == 1;
        const bool raw_in = l.KIND == IN_U8 || l.KIND == IN_U16;
        const bool split = split3 && first && l.NC == 3 && raw_in && l.MCT == MCTK_ICT;  // one component per job
        const bool x3 = split && use_x3;
        const int NP = (l.NC == 3 && !split) ? 2 : 4;
        const int SG = (raw_in && P.raw.sign_sub != 0) ? 1 : 0;
        const int ES = l.KIND == IN_U8 ? 1 : (l.KIND == IN_U16 ? 2 : 4);
        const int PB = ES * (raw_in ? l.NC : 1);
        if (a.px != 0 || a.hskip || a.vskip) return 0;
        bool seg_ua = false;           // this level needs the general-alignment variant
        if (a.w % (2 * NP)) seg_ua = true;  // aligned variant: whole-vector stores only (low and high band widths multiples of NP)
        if (first) {
            if (!ring_variant_supported(WT, NP, l.NC, l.KIND, l.MCT, SG)) return 0;
            if (l.NC == 1 && raw_in && s.C != 1) return 0;  // strided components
            if (!have_first) { R.WT = WT; R.NP1 = NP; R.NC1 = l.NC; R.IN1 = l.KIND; R.MCT1 = l.MCT; R.SG1 = SG; R.X3 = x3 ? 1 : 0; have_first = true; }
            else if (R.NP1 != NP || R.NC1 != l.NC || R.IN1 != l.KIND || R.MCT1 != l.MCT || R.SG1 != SG) return 0;
        } else {
            if (l.NC != 1 || l.KIND != (WT == 53 ? IN_I32 : IN_F32)) return 0;
        }
        // alignment of the staged side: row pitch, row length and every item origin are multiples of 16 bytes
        const long long pitch = (long long)a.x_row_stride * ES;
        const long long row_bytes = (long long)a.w * PB;
        if (row_bytes > 0x7fffffffLL) return 0;
        if (pitch % 16 || row_bytes % 16) seg_ua = true;
        int lalign = 16 | (int)(pitch & 15) | (PB == 1 ? 8 : 0);
        for (int i = 0; i < a.n_items; i++) {
            if ((tab[l.o_x + i] * ES) % 16) seg_ua = true;
            lalign |= (int)((tab[l.o_x + i] * ES) & 15);
        }
        // alignment of the band side for NP-wide vector stores
        if ((a.hl.row_stride % NP) || (a.lw % NP) || (a.hl.comp_stride % NP) || (a.ll.row_stride % NP) || (a.ll.comp_stride % NP) ||
            (a.hl.x_off % NP) || (a.hh.x_off % NP) || (a.ll.x_off % NP) || (a.lh_.x_off % NP))
            seg_ua = true;
        for (int i = 0; i < a.n_items; i++)
            if ((tab[l.o_plane + i] % NP) || (tab[l.o_ll + i] % NP)) seg_ua = true;
        if (seg_ua) {
            // the general-alignment variant exists for single-component jobs on raw words (level 1) and on LL planes (deeper
            // levels); windows narrower than one vector per lane pair stay on the per-level kernels
            if (!env_int("J2K_RING_UA", 1) || l.NC != 1 || NP != 4 || a.w < 16 || a.h < 2 || (first && !raw_in)) return 0;
            ua = true;
        }
        {   // widest access every lane address / band row is guaranteed to allow (powers of two; used by the UA variant only)
            g.load_align = lalign & -lalign;
            auto cls = [&](const BandIO& b, size_t o_tab) {
                long long m = 4 | b.row_stride | b.x_off;
                for (int i = 0; i < a.n_items; i++) m |= tab[o_tab + i];
                return (int)(m & -m);
            };
            g.st_cls[0] = cls(a.ll, l.o_ll); g.st_cls[1] = cls(a.hl, l.o_plane); g.st_cls[2] = cls(a.lh_, l.o_plane); g.st_cls[3] = cls(a.hh, l.o_plane);
            g.x_align = 16;
        }
        const BandIO* bands[4] = {&a.ll, &a.hl, &a.lh_, &a.hh};
        for (int bi = 0; bi < 4; bi++) {
            const BandIO& b = *bands[bi];
            FastQ& q = g.q[bi];
            q.mode = b.mode; q.shift = 0; q.step = 1.f; q.rcp = 1.f;
            if (WT == 53) { if (b.mode == Q_SHIFT) q.shift = b.shift; else if (b.mode != Q_RAW) return 0; }
            else if (b.mode == Q_QUANT) {
                if (b.rcp == 0.f) return 0;  // not a Markstein-safe step: the per-level kernels divide (div_by_step)
                q.step = b.step / b.scale; q.rcp = 1.0f / q.step;  // scale is a power of two: exact
            }
            else if (b.mode != Q_RAW) return 0;
        }
        g.rcpE = make_float2(g.q[0].rcp, g.q[2].rcp); g.nstE = make_float2(-g.q[0].step, -g.q[2].step);
        g.rcpO = make_float2(g.q[1].rcp, g.q[3].rcp); g.nstO = make_float2(-g.q[1].step, -g.q[3].step);
        g.w = a.w; g.h = a.h; g.py = a.py; g.lw = a.lw; g.lh = a.lh; g.Kx = a.Kx; g.Ky = a.Ky;
        g.n_items = split ? 3 * a.n_items : a.n_items;
        g.first = first ? 1 : 0;
        g.row_bytes = (int)row_bytes;
        g.dc = first ? a.raw.dc : 0;
        g.x_off = a.x_off;
        g.x_row_bytes = pitch;
        g.ll = a.ll; g.hl = a.hl; g.lh_ = a.lh_; g.hh = a.hh;
        if (x3) {
            // A pixel job of fwd3w_kernel occupies a quad (four warps), and a launch has only sms x 4 quads: the chunk height
            // is chosen so that the job count fills whole rounds of them.  Estimated makespan of n chunks of cp row pairs:
            // rounds x (cp + the 2 LAG warm-up pairs a chunk recomputes + the start-up bubble of a job, ~16 pairs measured).
            ring_chunks(g, NP, a.n_items, l.level, 2, env_int("J2K_FWD3W_TDIV", 4));
            const long long slots = (long long)device_sm_count() * J2K_F3_QUADS;
            // Small launches stay on the component jobs: with little more than one full-height job per quad the coarser levels no
            // longer overlap level 1 and the tail decides (C3 frames per launch, fwd3w against component jobs, with the job-size
            // policies below: 8: 0.586 / 0.622, 12: 0.671 / 0.651, 16: 0.713 / 0.672, 24: 0.732 / 0.689, 32: 0.74 / 0.70;
            // 128 C5 tiles: 0.655 / 0.60 - profiles/exp_r02_chunk_policy.log).  J2K_FWD3W=2 forces the kernel (tests).
            if (env_int("J2K_FWD3W", J2K_FWD3W_DEFAULT) < 2 &&
                (long long)a.n_items * g.nstrips * g.Ky < (long long)env_int("J2K_FWD3W_MIN_PAIRS", 160) * slots) return 0;
            if (!env_int("J2K_FWD3W_TDIV", 0)) {
                const int max_chunk = env_int("J2K_RING_CHUNK", 128), min_chunk = env_int("J2K_RING_CHUNK_MIN", 8);
                const int bubble = env_int("J2K_FWD3W_BUBBLE", 16);
                long long best = -1;
                for (int n = 1; n <= g.Ky; n++) {
                    const int cp = (g.Ky + n - 1) / n;
                    if (cp > max_chunk) continue;
                    if (cp < min_chunk && best >= 0) break;
                    const int nch = (g.Ky + cp - 1) / cp;
                    const long long nj = (long long)a.n_items * g.nstrips * nch;
                    const long long cost = ((nj + slots - 1) / slots) * (cp + 4 + bubble);
                    if (best < 0 || cost < best) { best = cost; g.chunk_pairs = cp; g.nchunks = nch; }
                }
            }
        } else {
            // Measured (profiles/exp_r02_chunk_policy.log): the three-components-per-lane 5/3 level 1 (RCT) runs best with
            // chunks of at most 32 row pairs (C3(ii) x 32 frames: 0.80 -> 0.88 of the HBM peak); the coarser levels behind
            // fwd3w_kernel, which only its twelve consumer warps per SM take, with 1.5 x the usual number of jobs
            // (C3(i) x 32: 0.71 -> 0.74, C5: 0.64 -> 0.65).
            const int cap = (first && WT == 53 && l.NC == 3) ? env_int("J2K_RING_CHUNK_RGB53", 32) : 0;
            ring_chunks(g, NP, g.n_items, l.level, 2, 1, (!first && R.X3) ? env_int("J2K_FWD3W_DEEP_HALVES", 3) : 2, cap);
        }
        g.dep_seg = -1; g.dep_div = 1; g.dep_target = 0;
        if (!first) {
            // producer: same class, previous level
            for (size_t pj = 0; pj < si; pj++) {
                const LevelLaunch& pl = P.levels[order[pj]];
                if (pl.cls == l.cls && pl.level == l.level - 1) {
                    g.dep_seg = (int)pj;
                    g.dep_div = (pl.nc3_first && R.NP1 == 2) ? 3 : 1;  // component-split producers count per component
                    g.dep_target = R.args.seg[pj].nchunks * R.args.seg[pj].nstrips;
                }
            }
            if (g.dep_seg < 0) return 0;
        }
        g.job_begin = jobs;
        long long nj = (long long)g.n_items * g.nchunks * g.nstrips;
        if (nj + jobs > 0x3fffffffLL) return 0;
        jobs += (int)nj;
        g.job_end = jobs;
        g.done_base = n_ctl;
        n_ctl += g.n_items;
        R.x_buf[si] = l.x_buf; R.ll_buf[si] = l.ll_buf; R.band_buf[si] = l.band_buf; R.level[si] = l.level;
        seg_of[order[si]] = (int)si;
    }
    if (!have_first) return 0;
    if (ua && (R.NC1 != 1 || R.NP1 != 4 || (R.IN1 != IN_U8 && R.IN1 != IN_U16))) return 0;  // a deep level needs the general variant but level 1 has none
    R.UA = ua ? 1 : 0;
    if (R.X3) {
        // jobs are claimed three at a time per CTA: every segment's range starts at a multiple of three (the padding numbers
        // decode to items past the end and are skipped); no pipelined slices for this variant
        int shift = 0;
        for (int k = 0; k < (int)order.size(); k++) {
            RingSeg& g = R.args.seg[k];
            g.job_begin += shift; g.job_end += shift;
            const int pad = (3 - g.job_end % 3) % 3;
            g.job_end += pad;
            shift += pad;
        }
        jobs += shift;
    }
    R.args.nseg = (int)order.size();
    for (int k = 0; k < R.args.nseg; k++) R.args.seg[k].has_waiters = 0;
    for (int k = 0; k < R.args.nseg; k++)
        if (R.args.seg[k].dep_seg >= 0) R.args.seg[R.args.seg[k].dep_seg].has_waiters = 1;
    R.args.total_jobs = jobs;
    R.args.n_ctl = n_ctl;
    R.args.raw = P.raw;
    R.args.one = 1.0f;
    int rc = ring_schedule(R, P, n_ctl);
    if (rc) return rc;
    R.ok = true;
    return 0;
}

int ring_blocks_per_sm(const void* fn, int smem_bytes) {
#ifdef J2K_EMU
    (void)fn; (void)smem_bytes;
    return 1;
#else
    int n = 0;
    if (smem_bytes > 48 * 1024 && cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) return -1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, J2K_RING_WARPS * 32, smem_bytes) != cudaSuccess) return -1;
    return n;
#endif
}

#define RING_CASE(wt, np, nc, in, mct, sg)                                                                                   \
    if (R.WT == wt && R.NP1 == np && R.NC1 == nc && R.IN1 == in && R.MCT1 == mct && R.SG1 == sg && !R.UA) {                  \
        if (query) return ring_blocks_per_sm((const void*)fwd_ring_kernel<wt, np, nc, in, mct, sg>, fwd_cta_smem<0>());      \
        J2K_LAUNCH_SMEM((fwd_ring_kernel<wt, np, nc, in, mct, sg>), grid, J2K_RING_WARPS * 32, fwd_cta_smem<0>(), st, A);    \
        return 0;                                                                                                           \
    }
#define RING_CASE_UA(wt, in, sg)                                                                                             \
    if (R.WT == wt && R.NP1 == 4 && R.NC1 == 1 && R.IN1 == in && R.MCT1 == MCTK_NONE && R.SG1 == sg && R.UA) {               \
        if (query) return ring_blocks_per_sm((const void*)fwd_ring_kernel<wt, 4, 1, in, MCTK_NONE, sg, 1>, fwd_cta_smem<1>()); \
        J2K_LAUNCH_SMEM((fwd_ring_kernel<wt, 4, 1, in, MCTK_NONE, sg, 1>), grid, J2K_RING_WARPS * 32, fwd_cta_smem<1>(), st, A); \
        return 0;                                                                                                           \
    }


// query = true: resident CTAs per SM of the variant (or -1); query = false: launch (0 = launched, -1 = no such variant)
#define RING_CASE_F3(in)                                                                                                     \
    if (R.IN1 == in) {                                                                                                      \
        if (query) {                                                                                                        \
            const void* fn = (const void*)fwd3w_kernel<in>;                                                                 \
            int n = 1;                                                                                                      \
            J2K_F3_QUERY(fn, n)                                                                                             \
            return n;                                                                                                       \
        }                                                                                                                   \
        J2K_LAUNCH_SMEM((fwd3w_kernel<in>), grid, J2K_F3_QUADS * 128, J2K_F3_CTA_SMEM, st, A);                                             \
        return 0;                                                                                                           \
    }
#ifdef J2K_EMU
#define J2K_F3_QUERY(fn, n) (void)fn;
#else
#define J2K_F3_QUERY(fn, n)                                                                                                  \
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, J2K_F3_CTA_SMEM) != cudaSuccess) return -1;    \
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, J2K_F3_QUADS * 128, J2K_F3_CTA_SMEM) != cudaSuccess) return -1;
#endif

int ring_dispatch_fwd(const RingPlan& R, const RingArgs& A, unsigned grid, cudaStream_t st, bool query) {
    if (R.X3) {
        RING_CASE_F3(IN_U8) RING_CASE_F3(IN_U16)
        return -1;
    }
    RING_CASE(97, 4, 1, IN_U8, MCTK_NONE, 0) RING_CASE(97, 4, 1, IN_U16, MCTK_NONE, 0)
    RING_CASE(97, 4, 1, IN_U8, MCTK_NONE, 1) RING_CASE(97, 4, 1, IN_U16, MCTK_NONE, 1)
    RING_CASE(97, 2, 3, IN_U8, MCTK_ICT, 0) RING_CASE(97, 2, 3, IN_U16, MCTK_ICT, 0)
    RING_CASE(97, 4, 3, IN_U8, MCTK_ICT, 0) RING_CASE(97, 4, 3, IN_U16, MCTK_ICT, 0)
    RING_CASE(97, 4, 1, IN_F32, MCTK_NONE, 0) RING_CASE(97, 4, 1, IN_I32, MCTK_NONE, 0)
    RING_CASE(53, 4, 1, IN_U8, MCTK_NONE, 0) RING_CASE(53, 4, 1, IN_U16, MCTK_NONE, 0)
    RING_CASE(53, 4, 1, IN_U8, MCTK_NONE, 1) RING_CASE(53, 4, 1, IN_U16, MCTK_NONE, 1)
    RING_CASE(53, 2, 3, IN_U8, MCTK_RCT, 0) RING_CASE(53, 2, 3, IN_U16, MCTK_RCT, 0)
    RING_CASE(53, 4, 1, IN_I32, MCTK_NONE, 0)
    RING_CASE_UA(97, IN_U8, 0) RING_CASE_UA(97, IN_U16, 0) RING_CASE_UA(97, IN_U8, 1) RING_CASE_UA(97, IN_U16, 1)
    RING_CASE_UA(53, IN_U8, 0) RING_CASE_UA(53, IN_U16, 0) RING_CASE_UA(53, IN_U8, 1) RING_CASE_UA(53, IN_U16, 1)
    return -1;
}


bool ring_inv_variant_supported(int WT, int NP, int NC, int OUT, int MCT) {
    if (NC == 3) return NP == (WT == 97 ? J2K_INV_RGB_NP : 2) && (OUT == IN_U8 || OUT == IN_U16) && MCT == (WT == 53 ? MCTK_RCT : MCTK_ICT);
    if (NP != 4 || MCT != MCTK_NONE) return false;
    if (OUT == IN_U8 || OUT == IN_U16) return true;
    return OUT == (WT == 53 ? IN_I32 : IN_F32);
}

// Inverse counterpart of build_ring_fwd: coarsest level first, every level waits for the level above it.
int build_ring_inv_impl(const Spec& s, Plan& P, const std::vector<long long>& tab, int cut);

// Levels cut..1 in the persistent launch (see build_ring_fwd); the coarser levels run first on the per-level kernels.
int build_ring_inv(const Spec& s, Plan& P, const std::vector<long long>& tab) {
    int maxlevel = 0;
    for (auto& l : P.levels) maxlevel = l.level > maxlevel ? l.level : maxlevel;
    for (int cut = maxlevel; cut >= 1; cut--) {
        int rc = build_ring_inv_impl(s, P, tab, cut);
        if (rc) return rc;
        if (P.ring.ok) {
            P.ring.cut = cut;
            for (auto& l : P.levels) l.in_ring = l.level <= cut;
            return 0;
        }
    }
    return 0;
}

int build_ring_inv_impl(const Spec& s, Plan& P, const std::vector<long long>& tab, int cut) {
    RingPlan& R = P.ring;
    R.ok = false;
    if (env_int("J2K_RING_DISABLE", 0) || env_int("J2K_RING_INV_DISABLE", 0)) return 0;
    if (P.levels.empty()) return 0;
    memset(&R.args, 0, sizeof R.args);
    const int WT = s.reversible ? 53 : 97;
    std::vector<int> order;
    for (int k = cut; k >= 1; k--)
        for (size_t i = 0; i < P.levels.size(); i++)
            if (P.levels[i].level == k) order.push_back((int)i);
    if (order.empty() || order.size() > J2K_RING_MAXSEG) return 0;
    bool have_first = false;
    bool ua = false, x3_any = false;
    int n_ctl = 2, jobs = 0;
    R.X3 = 0;
    // (known before the loop reaches level 1: the coarser levels of such a plan are cut into HALF the usual number of jobs -
    // measured on the final kernels, C3(i) x 32 inverse 0.602 -> 0.620, C5 0.513 -> 0.521; profiles/exp_r02_chunk_policy.log)
    bool x3_plan = false;
    for (int oi : order) {
        const LevelLaunch& l1 = P.levels[oi];
        if (l1.level == 1 && WT == 97 && l1.NC == 3 && l1.KIND == IN_U8 && l1.MCT == MCTK_ICT && env_int("J2K_INV3W", J2K_INV3W_DEFAULT) != 0) x3_plan = true;
    }
    for (size_t si = 0; si < order.size(); si++) {
        const LevelLaunch& l = P.levels[order[si]];
        const LevelArgs& a = l.a;
        RingSeg& g = R.args.seg[si];
        const bool first = l.level == 1;
        const bool raw_out = l.KIND == IN_U8 || l.KIND == IN_U16;
        // ICT + 9/7 on 8-bit RGB: level 1 as three single-component producers + one pixel consumer per CTA (inv3w_kernel)
        const bool x3 = first && WT == 97 && l.NC == 3 && l.KIND == IN_U8 && l.MCT == MCTK_ICT && env_int("J2K_INV3W", J2K_INV3W_DEFAULT) != 0;
        const int NP = l.NC == 3 ? (x3 ? 4 : (WT == 97 ? J2K_INV_RGB_NP : 2)) : 4;
        const int ES = l.KIND == IN_U8 ? 1 : (l.KIND == IN_U16 ? 2 : 4);
        const int PB = ES * (raw_out ? l.NC : 1);
        if (a.px != 0 || a.hskip || a.vskip) return 0;
        bool seg_ua = false;          // this level needs the general-alignment variant
        if (a.w % 8) seg_ua = true;   // aligned variant: band rows are staged in 16-byte units (low and high band widths multiples of 4)
        if (first) {
            if (!x3 && !ring_inv_variant_supported(WT, NP, l.NC, l.KIND, l.MCT)) return 0;
            if (l.NC == 1 && raw_out && s.C != 1) return 0;  // strided components
            if (!have_first) { R.WT = WT; R.NP1 = NP; R.NC1 = l.NC; R.IN1 = l.KIND; R.MCT1 = l.MCT; R.SG1 = 0; R.X3 = x3 ? 1 : 0; have_first = true; }
            else if (R.NP1 != NP || R.NC1 != l.NC || R.IN1 != l.KIND || R.MCT1 != l.MCT) return 0;
        } else {
            if (l.NC != 1 || l.KIND != (WT == 53 ? IN_I32 : IN_F32)) return 0;
        }
        // destination rows: vector stores
        const long long pitch = (long long)a.x_row_stride * ES;
        if (pitch % 16 || ((long long)a.w * PB) % 16) seg_ua = true;
        int xalign = 16 | (int)(pitch & 15);
        for (int i = 0; i < a.n_items; i++) {
            if ((tab[l.o_x + i] * ES) % 16) seg_ua = true;
            xalign |= (int)((tab[l.o_x + i] * ES) & 15);
        }
        // band rows: 16-byte aligned TMA segments
        if ((a.hl.row_stride % 4) || (a.lw % 4) || (a.hl.comp_stride % 4) || (a.ll.row_stride % 4) || (a.ll.comp_stride % 4) ||
            (a.hl.x_off % 4) || (a.hh.x_off % 4) || (a.ll.x_off % 4) || (a.lh_.x_off % 4))
            seg_ua = true;
        long long lal = 4 | a.hl.row_stride | a.ll.row_stride | a.hl.x_off | a.hh.x_off | a.ll.x_off | a.lh_.x_off;
        for (int i = 0; i < a.n_items; i++) {
            if ((tab[l.o_plane + i] % 4) || (tab[l.o_ll + i] % 4)) seg_ua = true;
            lal |= tab[l.o_plane + i] | tab[l.o_ll + i];
        }
        if (seg_ua) {
            if (!env_int("J2K_RING_UA", 1) || l.NC != 1 || NP != 4 || a.w < 16 || a.h < 2 || (first && !raw_out)) return 0;
            ua = true;
        }
        g.load_align = (int)(lal & -lal) * 4;   // bytes every staged lane address is aligned to (band samples are ints)
        g.x_align = xalign & -xalign;           // bytes the destination rows are aligned to
        g.st_cls[0] = g.st_cls[1] = g.st_cls[2] = g.st_cls[3] = 4;
        const BandIO* bands[4] = {&a.ll, &a.hl, &a.lh_, &a.hh};
        float scl[4];
        for (int bi = 0; bi < 4; bi++) {
            const BandIO& b = *bands[bi];
            scl[bi] = 1.f;
            if (WT == 53) { if (b.mode != DQ_RAW && b.mode != DQ_HALVE) return 0; }
            else if (b.mode == DQ_SCALE) scl[bi] = b.scale;
            else if (b.mode != DQ_RAW && b.mode != DQ_CVT) return 0;
        }
        g.rcpE = make_float2(scl[0], scl[2]); g.rcpO = make_float2(scl[1], scl[3]);
        g.nstE = g.nstO = make_float2(0.f, 0.f);
        g.w = a.w; g.h = a.h; g.py = a.py; g.lw = a.lw; g.lh = a.lh; g.Kx = a.Kx; g.Ky = a.Ky;
        g.n_items = a.n_items;
        g.first = first ? 1 : 0;
        g.row_bytes = (int)((long long)a.w * PB);
        g.dc = 0;
        g.x_off = a.x_off;
        g.x_row_bytes = pitch;
        g.x_mode = a.x_mode;
        g.ll = a.ll; g.hl = a.hl; g.lh_ = a.lh_; g.hh = a.hh;
        g.planes_out = nullptr; g.planes_off = a.planes_off; g.planes_comp_stride = a.planes_comp_stride; g.planes_row_stride = a.planes_row_stride;
        if (l.planes_buf) {
            long long pal = 4 | a.planes_row_stride | a.planes_comp_stride;
            for (int i = 0; i < a.n_items; i++)
                if (a.planes_off) pal |= tab[(size_t)(a.planes_off - (const long long*)P.tables.p) + i];
            if (pal & 3) {
                if (!env_int("J2K_RING_UA", 1) || l.NC != 1 || NP != 4 || a.w < 16 || a.h < 2) return 0;
                ua = true;
            }
            g.st_cls[0] = (int)(pal & -pal);  // GetImageData plane rows (UA)
        }
        ring_chunks(g, NP, g.n_items, l.level, (J2K_INV_HALO_FREE && WT == 53) ? 0 : 2, x3 ? env_int("J2K_INV3W_TDIV", 4) : 1,
                    (!first && x3_plan) ? env_int("J2K_INV3W_DEEP_HALVES", 1) : 2,
                    (first && WT == 53 && l.NC == 3) ? env_int("J2K_RING_CHUNK_RGB53", 32) : 0);   // (see build_ring_fwd_impl)
        // Single-component level 1 of the inverse, aligned variant: the chunk height that fills whole rounds of the resident
        // warps (level 1 is the END of an inverse launch, nothing has to overlap behind it; measured: C1 0.818 -> 0.840,
        // C4 0.816 -> 0.844, C2 x16 0.810 -> 0.818, C2 x32 unchanged; the forward, whose coarser levels must overlap its
        // level 1, loses with the same rule: C1 0.765 -> 0.670 - profiles/exp_r02_chunk_policy.log)
        if (first && l.NC == 1 && !seg_ua && env_int("J2K_RING_MAKESPAN", 1)) {
            const int warps = WT == 53 ? 4 * J2K_INV_MINB_53 : 4 * J2K_RING_MINB;
            ring_chunks_makespan(g, (long long)g.n_items * g.nstrips, (long long)device_sm_count() * warps, WT == 97 ? 4 : 2,
                                 env_int("J2K_RING_CHUNK", 128));
        }
        g.dep_seg = -1; g.dep_div = 1; g.dep_target = 0; g.dep_mul = 1;
        {
            // producer: same class, next coarser level (absent for the coarsest level of the class)
            for (size_t pj = 0; pj < si; pj++) {
                const LevelLaunch& pl = P.levels[order[pj]];
                if (pl.cls == l.cls && pl.level == l.level + 1) {
                    g.dep_seg = (int)pj;
                    g.dep_mul = l.nc3_first ? 3 : 1;  // a pixel item consumes the LL planes of its three components
                    g.dep_target = R.args.seg[pj].nchunks * R.args.seg[pj].nstrips;
                }
            }
        }
        g.job_begin = jobs;
        long long nj = (long long)g.n_items * g.nchunks * g.nstrips;
        if (x3) nj *= 3;  // a pixel job holds three job numbers, one per component (producer warp)
        if (nj + jobs > 0x3fffffffLL) return 0;
        jobs += (int)nj;
        g.job_end = jobs;
        x3_any = x3_any || x3;
        g.done_base = n_ctl;
        n_ctl += g.n_items;
        R.x_buf[si] = l.x_buf; R.ll_buf[si] = l.ll_buf; R.band_buf[si] = l.band_buf; R.level[si] = l.level;
        R.planes_buf[si] = l.planes_buf;
    }
    if (!have_first) return 0;
    if (ua && (R.NC1 != 1 || R.NP1 != 4 || (R.IN1 != IN_U8 && R.IN1 != IN_U16))) return 0;  // a coarse level needs the general variant but level 1 has none
    R.UA = ua ? 1 : 0;
    if (R.X3) {
        // jobs are claimed three at a time per CTA: every segment's range starts at a multiple of three (the padding numbers
        // decode to items past the end and are skipped); no pipelined slices for this variant
        if (order.size() > 1 && R.args.seg[0].first) return 0;
        int shift = 0;
        for (int k = 0; k < (int)order.size(); k++) {
            RingSeg& g = R.args.seg[k];
            g.job_begin += shift; g.job_end += shift;
            const int pad = (3 - g.job_end % 3) % 3;
            g.job_end += pad;   // padding numbers belong to the segment; they decode to an item past its end
            shift += pad;
        }
        jobs += shift;
    }
    R.args.nseg = (int)order.size();
    for (int k = 0; k < R.args.nseg; k++) R.args.seg[k].has_waiters = 0;
    for (int k = 0; k < R.args.nseg; k++)
        if (R.args.seg[k].dep_seg >= 0) R.args.seg[R.args.seg[k].dep_seg].has_waiters = 1;
    R.args.total_jobs = jobs;
    R.args.n_ctl = n_ctl;
    R.args.raw = P.raw;
    R.args.one = 1.0f;
    int rc = ring_schedule(R, P, n_ctl);
    if (rc) return rc;
    R.ok = true;
    return 0;
}

#define RING_INV_CASE_UA(wt, out)                                                                                            \
    if (R.WT == wt && R.NP1 == 4 && R.NC1 == 1 && R.IN1 == out && R.MCT1 == MCTK_NONE && R.UA) {                             \
        if (query) return ring_blocks_per_sm((const void*)inv_ring_kernel<wt, 4, 1, out, MCTK_NONE, 1>, (inv_cta_smem<wt, 1, 1>())); \
        J2K_LAUNCH_SMEM((inv_ring_kernel<wt, 4, 1, out, MCTK_NONE, 1>), grid, J2K_RING_WARPS * 32, (inv_cta_smem<wt, 1, 1>()), st, A); \
        return 0;                                                                                                           \
    }
#define RING_INV_CASE(wt, np, nc, out, mct)                                                                                  \
    if (R.WT == wt && R.NP1 == np && R.NC1 == nc && R.IN1 == out && R.MCT1 == mct && !R.UA) {                                \
        if (query) return ring_blocks_per_sm((const void*)inv_ring_kernel<wt, np, nc, out, mct>, (inv_cta_smem<wt, nc>()));   \
        J2K_LAUNCH_SMEM((inv_ring_kernel<wt, np, nc, out, mct>), grid, J2K_RING_WARPS * 32, (inv_cta_smem<wt, nc>()), st, A); \
        return 0;                                                                                                           \
    }

int ring_dispatch_inv(const RingPlan& R, const RingArgs& A, unsigned grid, cudaStream_t st, bool query) {
    if (R.X3) {
        if (query) {
#ifdef J2K_EMU
            return 1;
#else
            int n = 0;
            const void* fn = (const void*)inv3w_kernel<IN_U8>;
            if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, J2K_X3_CTA_SMEM) != cudaSuccess) return -1;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, 128, J2K_X3_CTA_SMEM) != cudaSuccess) return -1;
            return n;
#endif
        }
        J2K_LAUNCH_SMEM((inv3w_kernel<IN_U8>), grid, 128, J2K_X3_CTA_SMEM, st, A);
        return 0;
    }
    RING_INV_CASE(97, 4, 1, IN_U8, MCTK_NONE) RING_INV_CASE(97, 4, 1, IN_U16, MCTK_NONE)
    RING_INV_CASE(97, J2K_INV_RGB_NP, 3, IN_U8, MCTK_ICT) RING_INV_CASE(97, J2K_INV_RGB_NP, 3, IN_U16, MCTK_ICT)
    RING_INV_CASE(97, 4, 1, IN_F32, MCTK_NONE)
    RING_INV_CASE(53, 4, 1, IN_U8, MCTK_NONE) RING_INV_CASE(53, 4, 1, IN_U16, MCTK_NONE)
    RING_INV_CASE(53, 2, 3, IN_U8, MCTK_RCT) RING_INV_CASE(53, 2, 3, IN_U16, MCTK_RCT)
    RING_INV_CASE(53, 4, 1, IN_I32, MCTK_NONE)
    RING_INV_CASE_UA(97, IN_U8) RING_INV_CASE_UA(97, IN_U16) RING_INV_CASE_UA(53, IN_U8) RING_INV_CASE_UA(53, IN_U16)
    return -1;
}

// Builds the launch list of one direction.  Returns 0 or a negative status.
int build_plan(const Spec& s, int nframes, long long frame_samples, Plan& P) {
    P.fwd = s.fwd; P.nframes = nframes; P.C = s.C; P.W = s.W; P.H = s.H; P.L = s.L;
    P.bps = s.bit_depth <= 8 ? 1 : 2;
    P.reversible = s.reversible != 0;
    P.direct = s.direct;
    P.raw = make_raw(s);
    P.pix_kind = s.bit_depth <= 8 ? IN_U8 : IN_U16;
    P.frame_samples = frame_samples;
    P.coeffs_per_frame = (long long)s.W * s.H * s.C;
    P.want_planes = s.want_planes;
    if (!s.direct) { int rc = make_prog(s, P.prog); if (rc) return rc; }
    const int WT = s.reversible ? 53 : 97;
    const long long HW = (long long)s.W * s.H;

    // classes of tiles that share every level window
    std::vector<ClassTmp> classes;
    const int pm = (1 << (s.L > 0 ? (s.L > 20 ? 20 : s.L) : 0)) - 1;
    for (size_t t = 0; t < s.tiles.size(); t++) {
        const TileGeom& g = s.tiles[t];
        if (g.tw <= 0 || g.th <= 0) continue;
        size_t k = 0;
        for (; k < classes.size(); k++)
            if (classes[k].tw == g.tw && classes[k].th == g.th && (classes[k].ox & pm) == (g.ox & pm) && (classes[k].oy & pm) == (g.oy & pm)) break;
        if (k == classes.size()) {
            ClassTmp c;
            c.tw = g.tw; c.th = g.th; c.ox = g.ox; c.oy = g.oy;
            performed_levels(g.tw, g.th, g.ox, g.oy, s.L, c.lv, c.cw, c.ch);
            c.rects = go_band_rects(g.tw, g.th, g.ox, g.oy, s.L);
            c.regular = s.reversible || s.direct || geometry_regular(c.lv, c.cw, c.ch, c.rects, s.L);
            classes.push_back(c);
        }
        classes[k].tiles.push_back((int)t);
    }
    bool any_nolevel = false;
    for (auto& c : classes) if (c.lv.empty()) any_nolevel = true;
    const bool std_mct = s.mct_mode == J2K_MCT_NONE || s.mct_mode == J2K_MCT_RCT || s.mct_mode == J2K_MCT_ICT;
    P.generic = !s.direct && (!std_mct || s.planar_in || any_nolevel);
    const bool nc3 = !P.generic && !s.direct && (s.mct_mode == J2K_MCT_RCT || s.mct_mode == J2K_MCT_ICT);
    const bool temp_is_f32 = s.fwd && s.mct_mode == J2K_MCT_ICT;  // encoder.go:206

    TableBuilder tb;
    struct Fix { int which; size_t level; size_t off; };  // which: 0 x_off, 1 ll, 2 hl, 3 lh, 4 hh, 5 planes
    std::vector<Fix> fixes;
    struct PwFix { bool post; size_t i; size_t src, dst; };
    std::vector<PwFix> pwfixes;
    long long sa_total = 0, sb_total = 0;

    auto pad4 = [](long long v) { return (v + 3) & ~3LL; };

    for (auto& c : classes) {
        const int nt = (int)c.tiles.size();
        const int Lp = (int)c.lv.size();
        const long long tn = (long long)c.tw * c.th;
        const long long sa_item = Lp >= 1 ? pad4((long long)c.lv[0].lw * c.lv[0].lh) : 0;
        const long long sb_item = Lp >= 2 ? pad4((long long)c.lv[1].lw * c.lv[1].lh) : 0;
        const long long n_comp_items = (long long)nframes * nt * s.C;
        const long long n_pix_items = (long long)nframes * nt;
        // offset tables (elements)
        std::vector<long long> plane_c(n_comp_items), sa_c(n_comp_items), sb_c(n_comp_items), x_c(n_comp_items), img_c(n_comp_items);
        std::vector<long long> plane_p(n_pix_items), sa_p(n_pix_items), x_p(n_pix_items), img_p(n_pix_items);
        for (int f = 0; f < nframes; f++)
            for (int ti = 0; ti < nt; ti++) {
                const TileGeom& g = s.tiles[c.tiles[ti]];
                long long pi = (long long)f * nt + ti;
                long long tile_pix = (long long)g.y0 * s.W + g.x0;
                plane_p[pi] = (long long)f * P.coeffs_per_frame + g.coeff_off;
                sa_p[pi] = sa_total + pi * s.C * sa_item;
                x_p[pi] = (long long)f * frame_samples + tile_pix * s.C;          // interleaved pixels
                img_p[pi] = (long long)f * s.C * HW + tile_pix;                    // planar image planes (temp / planes_out)
                for (int comp = 0; comp < s.C; comp++) {
                    long long ci = pi * s.C + comp;
                    plane_c[ci] = plane_p[pi] + comp * tn;
                    sa_c[ci] = sa_p[pi] + comp * sa_item;
                    sb_c[ci] = sb_total + ci * sb_item;
                    x_c[ci] = x_p[pi] + comp;
                    img_c[ci] = img_p[pi] + comp * HW;
                }
            }
        const size_t o_plane_c = tb.add(plane_c), o_sa_c = tb.add(sa_c), o_sb_c = tb.add(sb_c), o_x_c = tb.add(x_c), o_img_c = tb.add(img_c);
        const size_t o_plane_p = tb.add(plane_p), o_sa_p = tb.add(sa_p), o_x_p = tb.add(x_p), o_img_p = tb.add(img_p);
        sa_total += n_comp_items * sa_item;
        sb_total += n_comp_items * sb_item;

        RectTable rt;
        fill_rect_table(s, c.rects, rt);
        const bool irregular = !c.regular;
        // which buffer holds the coefficient planes the level kernels touch
        const int coef_buf = (!s.fwd && irregular) ? B_CTEMP : B_COEF;

        if (Lp == 0) {
            // no level is performed (num_levels == 0 or <= 1x1 tiles): copy windows + literal band passes
            PwLaunch pw{};
            pw.n_items = (int)n_comp_items; pw.w = c.tw; pw.h = c.th;
            if (s.fwd) {
                if (s.direct) continue;  // nothing to do: data is unchanged
                pw.kind = PW_COPY; pw.src_buf = B_TEMP; pw.dst_buf = B_COEF; pw.src_stride = s.W; pw.dst_stride = c.tw;
                pw.cvt = (s.reversible || temp_is_f32) ? 0 : 1;
                pwfixes.push_back({true, P.post.size(), o_img_c, o_plane_c});
                P.post.push_back(pw);
                if (!s.reversible) {
                    PwLaunch q{};
                    q.kind = PW_QUANT_RECTS; q.dst_buf = B_COEF; q.n_items = (int)n_comp_items; q.w = c.tw; q.h = c.th; q.dst_stride = c.tw;
                    q.all_round = (s.L == 0 || s.n_steps == 0); q.rt = rt;
                    pwfixes.push_back({true, P.post.size(), o_plane_c, o_plane_c});
                    P.post.push_back(q);
                } else if (s.fuse_shift && !s.htj2k) {
                    PwLaunch q{};
                    q.kind = PW_SHIFT; q.dst_buf = B_COEF; q.n_items = (int)n_comp_items; q.w = c.tw; q.h = c.th; q.dst_stride = c.tw; q.shift = 6;
                    pwfixes.push_back({true, P.post.size(), o_plane_c, o_plane_c});
                    P.post.push_back(q);
                }
            } else {
                if (s.direct) continue;
                if (s.reversible || s.L == 0) {  // t2/tile_decoder.go:887-898
                    pw.kind = PW_COPY; pw.src_buf = B_COEF; pw.dst_buf = B_TEMP; pw.src_stride = c.tw; pw.dst_stride = s.W;
                    pw.cvt = (s.reversible && s.fuse_shift && !s.htj2k) ? 3 : 0;
                    pwfixes.push_back({false, P.pre.size(), o_plane_c, o_img_c});
                    P.pre.push_back(pw);
                } else {                          // dequantize, (no-op inverse), round: t2/tile_decoder.go:899-913
                    pw.kind = PW_DEQUANT_RECTS; pw.src_buf = B_COEF; pw.dst_buf = B_TEMP; pw.src_stride = c.tw; pw.dst_stride = s.W; pw.rt = rt;
                    pwfixes.push_back({false, P.pre.size(), o_plane_c, o_img_c});
                    P.pre.push_back(pw);
                    PwLaunch r2{};
                    r2.kind = PW_COPY; r2.src_buf = B_TEMP; r2.dst_buf = B_TEMP; r2.src_stride = s.W; r2.dst_stride = s.W;
                    r2.n_items = (int)n_comp_items; r2.w = c.tw; r2.h = c.th; r2.cvt = 2;
                    pwfixes.push_back({false, P.pre.size(), o_img_c, o_img_c});
                    P.pre.push_back(r2);
                }
            }
            continue;
        }

        if (!s.fwd && irregular) {  // literal dequantization into a float copy of the planes
            PwLaunch pw{};
            pw.kind = PW_DEQUANT_RECTS; pw.src_buf = B_COEF; pw.dst_buf = B_CTEMP; pw.src_stride = c.tw; pw.dst_stride = c.tw;
            pw.n_items = (int)n_comp_items; pw.w = c.tw; pw.h = c.th; pw.rt = rt;
            pwfixes.push_back({false, P.pre.size(), o_plane_c, o_plane_c});
            P.pre.push_back(pw);
        }

        // level launches: forward k = 1..Lp, inverse k = Lp..1
        for (int step_i = 0; step_i < Lp; step_i++) {
            const int k = s.fwd ? step_i + 1 : Lp - step_i;  // 1-based level
            const LevelGeom& g = c.lv[k - 1];
            LevelLaunch l{};
            LevelArgs& a = l.a;
            memset(&a, 0, sizeof a);
            set_geom(a, g);
            l.WT = WT;
            l.level = k;
            a.raw = P.raw;
            const bool first = (k == 1);       // touches the image side
            const bool last = (k == Lp);       // touches the coarsest LL
            const bool use_nc3 = nc3 && first;
            l.NC = use_nc3 ? 3 : 1;
            l.MCT = use_nc3 ? (s.mct_mode == J2K_MCT_RCT ? MCTK_RCT : MCTK_ICT) : MCTK_NONE;
            a.n_items = (int)(use_nc3 ? n_pix_items : n_comp_items);
            const int res = s.L - (k - 1);     // resolution whose bands this level holds
            const int bidx = 3 * (res - 1) + 1;

            // ---- interleaved side
            size_t o_x;
            if (first) {
                if (s.direct) {
                    l.KIND = s.reversible ? IN_I32 : IN_F32;
                    l.x_buf = s.fwd ? B_PIX : B_PLANES; l.x_elem_bytes = 4;
                    a.x_row_stride = s.W; a.x_comp_stride = HW; o_x = o_img_c;
                    a.raw.dc = 0; a.x_mode = 0;
                } else if (P.generic) {
                    l.KIND = s.fwd ? (temp_is_f32 ? IN_F32 : IN_I32) : (s.reversible ? IN_I32 : IN_F32);
                    l.x_buf = B_TEMP; l.x_elem_bytes = 4;
                    a.x_row_stride = s.W; a.x_comp_stride = HW; o_x = o_img_c;
                    a.raw.dc = 0;
                    a.x_mode = 1;  // inverse: store rounded int32 samples (finalize_kernel does the rest)
                } else {
                    l.KIND = P.pix_kind;
                    l.x_buf = B_PIX; l.x_elem_bytes = P.bps;
                    a.x_row_stride = s.W * s.C; o_x = use_nc3 ? o_x_p : o_x_c;
                    a.x_mode = 1;
                    if (!s.fwd && s.want_planes) {
                        l.planes_buf = B_PLANES;
                        a.planes_comp_stride = HW; a.planes_row_stride = s.W;
                        fixes.push_back({5, P.levels.size(), use_nc3 ? o_img_p : o_img_c});
                    }
                }
            } else {
                l.KIND = s.reversible ? IN_I32 : IN_F32;
                // forward reads LL_{k-1}; inverse writes LL_{k-1}
                l.x_buf = ((k - 1) & 1) ? B_SA : B_SB; l.x_elem_bytes = 4;
                a.x_row_stride = c.lv[k - 2].lw; o_x = ((k - 1) & 1) ? o_sa_c : o_sb_c;
                a.raw.dc = 0; a.x_mode = 0;
            }
            a.x_kind = l.KIND;
            fixes.push_back({0, P.levels.size(), o_x});

            // ---- band side
            auto band = [&](BandIO& b, int xo, int yo, int idx) {
                b.row_stride = c.tw; b.x_off = xo; b.y_off = yo; b.comp_stride = tn;
                if (s.fwd) band_mode_fwd(s, idx, b); else band_mode_inv(s, idx, b);
                if (s.fwd && irregular) b.mode = Q_RAW;         // literal quantization runs afterwards
                if (!s.fwd && irregular) b.mode = DQ_RAW;       // already dequantized into B_CTEMP
            };
            band(a.hl, g.lw, 0, bidx);
            band(a.lh_, 0, g.lh, bidx + 1);
            band(a.hh, g.lw, g.lh, bidx + 2);
            l.band_buf = coef_buf;
            const size_t o_plane = use_nc3 ? o_plane_p : o_plane_c;
            fixes.push_back({2, P.levels.size(), o_plane});
            fixes.push_back({3, P.levels.size(), o_plane});
            fixes.push_back({4, P.levels.size(), o_plane});
            size_t o_ll;
            if (last) {
                band(a.ll, 0, 0, 0);
                l.ll_buf = coef_buf;
                o_ll = o_plane;
            } else {
                a.ll.row_stride = g.lw; a.ll.x_off = 0; a.ll.y_off = 0;
                a.ll.mode = s.fwd ? (int)Q_RAW : (int)DQ_RAW; a.ll.step = a.ll.rcp = a.ll.scale = 1.f;
                a.ll.comp_stride = (k & 1) ? sa_item : sb_item;
                l.ll_buf = (k & 1) ? B_SA : B_SB;
                o_ll = (k & 1) ? (use_nc3 ? o_sa_p : o_sa_c) : o_sb_c;
            }
            fixes.push_back({1, P.levels.size(), o_ll});

            // ---- kernel shape
            if (l.NC == 3) l.NP = 2;
            else if (l.KIND == IN_U8 || l.KIND == IN_U16) l.NP = (a.Kx >= 64) ? 4 : 1;
            else l.NP = (a.Kx > valid_pairs_of(WT, 1)) ? 2 : 1;
            choose_chunks(a, valid_pairs_of(WT, l.NP));

            // ---- 128-bit fast paths: every offset, stride and band origin a multiple of the vector width
            const int NS = 2 * l.NP;
            {
                const std::vector<long long>& H = tb.host;
                size_t n = (size_t)a.n_items;
                bool okx = g.px == 0 && (a.x_row_stride % NS) == 0 && all_mult(H, o_x, o_x + n, NS);
                if (l.NC == 3 && first && !P.generic) okx = g.px == 0 && (a.x_row_stride % 4) == 0 && all_mult(H, o_x, o_x + n, 4);
                if (l.NC == 1 && first && !P.generic && !s.direct && s.C != 1) okx = false;  // strided components
                a.vec_x = okx;
                bool okb = g.px == 0 && (c.tw % l.NP) == 0 && (g.lw % l.NP) == 0 && (tn % l.NP) == 0 &&
                           all_mult(H, o_plane, o_plane + n, l.NP) && (a.ll.row_stride % l.NP) == 0 &&
                           (a.ll.comp_stride % l.NP) == 0 && all_mult(H, o_ll, o_ll + n, l.NP);
                a.vec_b = okb;
            }
            l.cls = (int)(&c - &classes[0]);
            l.nc3_first = use_nc3;
            l.o_x = o_x; l.o_ll = o_ll; l.o_plane = o_plane;
            l.fast = false;
            if (s.fwd && a.vec_x && a.vec_b && g.px == 0 && !a.hskip && !a.vskip && has_fwd_fast(l)) {
                const BandIO* bands[4] = {&a.ll, &a.hl, &a.lh_, &a.hh};
                bool ok = true;
                for (int bi = 0; bi < 4; bi++) {
                    const BandIO& b = *bands[bi];
                    FastQ& q = l.fq.q[bi];
                    q.mode = b.mode; q.shift = 0; q.step = 1.f; q.rcp = 1.f;
                    if (WT == 53) { if (b.mode == Q_SHIFT) q.shift = b.shift; else if (b.mode != Q_RAW) ok = false; }
                    else if (b.mode == Q_QUANT) {
                        if (b.rcp == 0.f) ok = false;  // not a Markstein-safe step: the generic kernel divides
                        q.step = b.step / b.scale; q.rcp = 1.0f / q.step;  // scale is a power of two: exact
                    }
                    else if (b.mode != Q_RAW) ok = false;
                }
                l.fast = ok;
            }
            P.levels.push_back(l);
        }

        if (s.fwd && irregular && !s.direct) {
            PwLaunch q{};
            q.kind = PW_QUANT_RECTS; q.dst_buf = B_COEF; q.n_items = (int)n_comp_items; q.w = c.tw; q.h = c.th; q.dst_stride = c.tw;
            q.all_round = (s.n_steps == 0); q.rt = rt;
            pwfixes.push_back({true, P.post.size(), o_plane_c, o_plane_c});
            P.post.push_back(q);
        }
    }

    // ---- device allocations and table upload
    int rc;
    if ((rc = P.tables.ensure(tb.host.size() * sizeof(long long) + 16))) return rc;
    CK(cudaMemcpy(P.tables.p, tb.host.data(), tb.host.size() * sizeof(long long), cudaMemcpyHostToDevice));
    if (sa_total && (rc = P.sa.ensure((size_t)sa_total * 4))) return rc;
    if (sb_total && (rc = P.sb.ensure((size_t)sb_total * 4))) return rc;
    if (P.generic && (rc = P.temp.ensure((size_t)nframes * s.C * HW * 4))) return rc;
    bool need_ctemp = false;
    for (auto& l : P.levels) if (l.band_buf == B_CTEMP) need_ctemp = true;
    if (need_ctemp && (rc = P.ctemp.ensure((size_t)nframes * P.coeffs_per_frame * 4))) return rc;
    const long long* T = (const long long*)P.tables.p;
    for (auto& f : fixes) {
        LevelArgs& a = P.levels[f.level].a;
        switch (f.which) {
        case 0: a.x_off = T + f.off; break;
        case 1: a.ll.off = T + f.off; break;
        case 2: a.hl.off = T + f.off; break;
        case 3: a.lh_.off = T + f.off; break;
        case 4: a.hh.off = T + f.off; break;
        case 5: a.planes_off = T + f.off; break;
        }
    }
    for (auto& f : pwfixes) {
        PwLaunch& pw = f.post ? P.post[f.i] : P.pre[f.i];
        pw.src_off = T + f.src; pw.dst_off = T + f.dst;
    }
    P.launches_per_run = (int)(P.levels.size() + P.pre.size() + P.post.size()) + (P.generic ? 1 : 0);
    {
        int rc2 = s.fwd ? build_ring_fwd(s, P, tb.host) : build_ring_inv(s, P, tb.host);
        if (rc2) return rc2;
    }
    return 0;
}

// ------------------------------------------------------------------ running a plan

inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

int run_pw(const PwLaunch& pw, void* const* bufs, cudaStream_t st) {
    long long total = (long long)pw.w * pw.h * pw.n_items;
    if (total <= 0) return 0;
    unsigned grid = (unsigned)((total + 255) / 256);
    switch (pw.kind) {
    case PW_COPY:
        J2K_LAUNCH(copy_window_kernel, grid, 256, st, (const int*)bufs[pw.src_buf], pw.src_off, pw.src_stride, (int*)bufs[pw.dst_buf],
                   pw.dst_off, pw.dst_stride, pw.n_items, pw.w, pw.h, pw.cvt);
        break;
    case PW_QUANT_RECTS:
        J2K_LAUNCH(quant_rects_kernel, grid, 256, st, (int*)bufs[pw.dst_buf], pw.dst_off, pw.n_items, pw.w, pw.h, pw.dst_stride, pw.rt,
                   pw.all_round);
        break;
    case PW_SHIFT:
        J2K_LAUNCH(shift_planes_kernel, grid, 256, st, (int*)bufs[pw.dst_buf], pw.dst_off, pw.n_items, pw.w, pw.h, pw.dst_stride, pw.shift);
        break;
    case PW_DEQUANT_RECTS:
        J2K_LAUNCH(dequant_rects_kernel, grid, 256, st, (const int*)bufs[pw.src_buf], pw.src_off, pw.src_stride, (int*)bufs[pw.dst_buf],
                   pw.dst_off, pw.dst_stride, pw.n_items, pw.w, pw.h, pw.rt);
        break;
    default: return fail(J2K_ERR_UNSUPPORTED, "internal: bad pointwise kind");
    }
    CK(cudaGetLastError());
    return 0;
}

// pixels / coeffs / planes are device pointers.  Forward: pixels -> coeffs.  Inverse: coeffs -> pixels (+planes).
// For the planar forward entry `pixels` holds nframes x C x H x W int32.  For direct plans `pixels` is the
// forward source plane and `planes` the inverse destination plane.
struct ProfScope {
    j2k_ctx* ctx; cudaStream_t st; bool on;
    ProfScope(j2k_ctx* c, cudaStream_t s) : ctx(c), st(s), on(c->profiling) { if (on) { ctx->prof_level.clear(); ctx->prof_stream = s; } }
    void begin(int level) {
        if (!on) return;
        size_t i = ctx->prof_level.size();
        while (ctx->prof_ev.size() < 2 * (i + 1)) { cudaEvent_t e; cudaEventCreate(&e); ctx->prof_ev.push_back(e); }
        ctx->prof_level.push_back(level);
        cudaEventRecord(ctx->prof_ev[2 * i], st);
    }
    void end() { if (on) cudaEventRecord(ctx->prof_ev[2 * (ctx->prof_level.size() - 1) + 1], st); }
};

int run_plan(j2k_ctx* ctx, Plan& P, void* pixels, void* coeffs, void* planes, bool planar_in, cudaStream_t st) {
    ProfScope prof(ctx, st);
    void* bufs[B_COUNT] = {nullptr};
    bufs[B_PIX] = pixels; bufs[B_COEF] = coeffs; bufs[B_PLANES] = planes;
    bufs[B_TEMP] = P.temp.p; bufs[B_SA] = P.sa.p; bufs[B_SB] = P.sb.p; bufs[B_CTEMP] = P.ctemp.p;
    const long long HW = (long long)P.W * P.H;
    int nl = 0;
    if (P.fwd && P.generic) {
        long long total = HW * P.nframes;
        unsigned grid = (unsigned)((total + 255) / 256);
        prof.begin(0);
        J2K_LAUNCH(prep_kernel, grid, 256, st, pixels, planar_in ? (int)IN_I32 : P.pix_kind,
                   planar_in ? (long long)P.C * HW : P.frame_samples, HW, P.C, P.raw, P.prog, (int*)P.temp.p, (long long)P.C * HW, P.nframes);
        prof.end();
        CK(cudaGetLastError());
        nl++;
    }
    for (auto& pw : P.pre) { prof.begin(0); int rc = run_pw(pw, bufs, st); prof.end(); if (rc) return rc; nl++; }
    bool use_ring = P.ring.ok;
    // per-level kernels: every level when the persistent launch is off, else the levels beyond its cut
    auto run_levels = [&]() -> int {
        for (auto& l : P.levels) {
            if (use_ring && l.in_ring) continue;
            LevelArgs a = l.a;
            a.x_base = bufs[l.x_buf];
            a.ll.base = bufs[l.ll_buf];
            a.hl.base = a.lh_.base = a.hh.base = bufs[l.band_buf];
            a.planes_out = l.planes_buf ? (int32_t*)bufs[l.planes_buf] : nullptr;
            if (!aligned16(a.x_base)) a.vec_x = 0;
            if (!aligned16(a.ll.base) || !aligned16(a.hl.base)) a.vec_b = 0;
            static const bool trace = getenv("J2K_B200_TRACE") != nullptr;
            if (trace)
                fprintf(stderr, "[j2k] %s level %d %dx%d WT=%d NP=%d NC=%d kind=%d mct=%d items=%d chunks=%d strips=%d vec_x=%d vec_b=%d fast=%d\n",
                        P.fwd ? "fwd" : "inv", l.level, a.w, a.h, l.WT, l.NP, l.NC, l.KIND, l.MCT, a.n_items, a.nchunks, a.nstrips, a.vec_x, a.vec_b,
                        (int)(l.fast && a.vec_x && a.vec_b));
            prof.begin(l.level);
            cudaError_t e = P.fwd ? ((l.fast && a.vec_x && a.vec_b) ? launch_fwd_fast(l, a, st) : launch_fwd_level(l, a, st)) : launch_inv_level(l, a, st);
            prof.end();
            if (e != cudaSuccess)
                return fail(J2K_ERR_CUDA, "level kernel launch (WT=%d NP=%d NC=%d kind=%d mct=%d) failed: %s", l.WT, l.NP, l.NC, l.KIND, l.MCT,
                            cudaGetErrorString(e));
            nl++;
        }
        return 0;
    };
    if (use_ring) {
        for (int i = 0; i < P.ring.args.nseg && use_ring; i++)
            use_ring = aligned16(bufs[P.ring.x_buf[i]]) && aligned16(bufs[P.ring.ll_buf[i]]) && aligned16(bufs[P.ring.band_buf[i]]) &&
                       (P.fwd || !P.ring.planes_buf[i] || aligned16(bufs[P.ring.planes_buf[i]]));
    }
    if (!P.fwd || !use_ring) {  // inverse: the coarse levels beyond the cut come first; no persistent launch: every level
        int rc = run_levels();
        if (rc) return rc;
    }
    if (use_ring) {
        RingPlan& R = P.ring;
        if (R.grid == 0) {
            int per_sm = P.fwd ? ring_dispatch_fwd(R, R.args, 0, st, true) : ring_dispatch_inv(R, R.args, 0, st, true);
            if (per_sm <= 0) return fail(J2K_ERR_CUDA, "ring kernel variant (WT=%d NP=%d NC=%d kind=%d mct=%d sg=%d) cannot be resident", R.WT, R.NP1,
                                         R.NC1, R.IN1, R.MCT1, R.SG1);
            int sms = 1;
#ifndef J2K_EMU
            int devid = 0;
            CK(cudaGetDevice(&devid));
            CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, devid));
#endif
            long long want = (long long)sms * per_sm;
            int cap = env_int("J2K_RING_GRID", 0);
            if (cap > 0 && cap < want) want = cap;
            R.grid = (unsigned)want;
        }
        RingArgs A = R.args;
        for (int i = 0; i < A.nseg; i++) {
            RingSeg& g = A.seg[i];
            g.x_base = (const unsigned char*)bufs[R.x_buf[i]];
            g.ll.base = bufs[R.ll_buf[i]];
            g.hl.base = g.lh_.base = g.hh.base = bufs[R.band_buf[i]];
            g.planes_out = (!P.fwd && R.planes_buf[i]) ? (int32_t*)bufs[R.planes_buf[i]] : nullptr;
        }
        static const bool trace = getenv("J2K_B200_TRACE") != nullptr;
        static const bool per_level = env_int("J2K_RING_PER_LEVEL", 0) != 0;
        static const bool dry = env_int("J2K_RING_DRY", 0) != 0;  // diagnostic: launch overhead only (no job is claimed)
        if (dry) A.total_jobs = 0;
        if (trace)
            fprintf(stderr, "[j2k] %s ring WT=%d NP=%d NC=%d kind=%d mct=%d sg=%d segs=%d jobs=%d slices=%d grid=%u x3=%d\n", P.fwd ? "fwd" : "inv", R.WT, R.NP1, R.NC1,
                    R.IN1, R.MCT1, R.SG1, A.nseg, A.total_jobs, A.nslice, R.grid, R.X3);
        if (!per_level) {
            unsigned grid = R.grid;
            // (three-producer inverse: a CTA per job triple; one-producer forward: four quads per CTA, a triple each)
            unsigned need = R.X3 ? (unsigned)((A.total_jobs + 2) / 3) : (unsigned)((A.total_jobs + J2K_RING_WARPS - 1) / J2K_RING_WARPS);
            if (R.X3 && P.fwd) need = (need + J2K_F3_QUADS - 1) / J2K_F3_QUADS;
            if (need < grid) grid = need;
            if (grid < 1) grid = 1;
            prof.begin(100);
            int rc = P.fwd ? ring_dispatch_fwd(R, A, grid, st, false) : ring_dispatch_inv(R, A, grid, st, false);
            prof.end();
            if (rc) return fail(J2K_ERR_CUDA, "ring kernel variant missing");
            CK(cudaGetLastError());
            nl++;
        } else {
            // diagnostic mode: one launch per level (same kernel, dependencies satisfied by stream order)
            int i = 0;
            while (i < A.nseg) {
                int j = i;
                while (j < A.nseg && R.level[j] == R.level[i]) j++;
                RingArgs B = A;
                B.nseg = j - i;
                B.nslice = 0;
                int base = A.seg[i].job_begin;
                for (int k = i; k < j; k++) {
                    B.seg[k - i] = A.seg[k];
                    B.seg[k - i].job_begin -= base; B.seg[k - i].job_end -= base;
                    B.seg[k - i].dep_seg = -1;
                    B.seg[k - i].has_waiters = 0;
                }
                B.total_jobs = A.seg[j - 1].job_end - base;
                unsigned grid = R.grid;
                unsigned need = (unsigned)((B.total_jobs + J2K_RING_WARPS - 1) / J2K_RING_WARPS);
                if (need < grid) grid = need;
                prof.begin(R.level[i]);
                int rc = P.fwd ? ring_dispatch_fwd(R, B, grid, st, false) : ring_dispatch_inv(R, B, grid, st, false);
                prof.end();
                if (rc) return fail(J2K_ERR_CUDA, "ring kernel variant missing");
                CK(cudaGetLastError());
                nl++;
                i = j;
            }
        }
    }
    if (P.fwd && use_ring) {  // forward: the levels beyond the cut follow the persistent launch in stream order
        int rc = run_levels();
        if (rc) return rc;
    }
    for (auto& pw : P.post) { prof.begin(0); int rc = run_pw(pw, bufs, st); prof.end(); if (rc) return rc; nl++; }
    if (!P.fwd && P.generic) {
        long long total = HW * P.nframes;
        unsigned grid = (unsigned)((total + 255) / 256);
        prof.begin(0);
        J2K_LAUNCH(finalize_kernel, grid, 256, st, (const int*)P.temp.p, (long long)P.C * HW, HW, P.C, P.raw, P.prog, pixels, P.pix_kind,
                   P.frame_samples, (int*)planes, (long long)P.C * HW, P.nframes);
        prof.end();
        CK(cudaGetLastError());
        nl++;
    }
    ctx->launches += nl;
    return nl;
}

// ------------------------------------------------------------------ parameter validation -> Spec

int validate_common(int w, int h, int C, int B, int L) {
    if (w <= 0 || h <= 0) return fail(J2K_ERR_INVALID_ARG, "invalid dimensions: %dx%d", w, h);                                  // encoder.go:294-296
    if (C <= 0 || C > J2K_MAX_COMPONENTS) return fail(J2K_ERR_INVALID_ARG, "invalid number of components: %d (must be 1-4)", C); // encoder.go:298-300
    if (B < 1 || B > 16) return fail(J2K_ERR_INVALID_ARG, "invalid bit depth: %d (must be 1-16)", B);                            // encoder.go:302-304
    if (L < 0 || L > J2K_MAX_LEVELS) return fail(J2K_ERR_INVALID_ARG, "invalid decomposition levels: %d (must be 0-%d)", L, J2K_MAX_LEVELS);
    if ((long long)w * h * C > (1LL << 34)) return fail(J2K_ERR_INVALID_ARG, "frame too large");  // frame-level offsets are 64-bit; a tile stays below 2^31 samples (checked per tile)
    return 0;
}

int fwd_tiles(const j2k_fwd_params* p, std::vector<TileGeom>* out) {  // encoder.go:1966-1997
    int tw = p->tile_width ? p->tile_width : p->width, th = p->tile_height ? p->tile_height : p->height;
    if (tw <= 0 || th <= 0) return 0;
    int ntx = (p->width + tw - 1) / tw, nty = (p->height + th - 1) / th;
    if (out) {
        long long off = 0;
        for (int t = 0; t < ntx * nty; t++) {
            int tx = t % ntx, ty = t / ntx;
            TileGeom g;
            g.x0 = tx * tw; g.y0 = ty * th;
            int x1 = g.x0 + tw > p->width ? p->width : g.x0 + tw, y1 = g.y0 + th > p->height ? p->height : g.y0 + th;
            g.tw = x1 - g.x0; g.th = y1 - g.y0; g.ox = g.x0; g.oy = g.y0; g.coeff_off = off;
            off += (long long)g.tw * g.th * p->components;
            out->push_back(g);
        }
    }
    return ntx * nty;
}

int inv_tiles(const j2k_inv_params* p, std::vector<TileGeom>* out) {  // tile_assembler.go:33-101, t2/tile_decoder.go:269-294
    int ntx = ceil_div(p->xsiz - p->xtosiz, p->xtsiz), nty = ceil_div(p->ysiz - p->ytosiz, p->ytsiz);
    if (out) {
        long long off = 0;
        for (int t = 0; t < ntx * nty; t++) {
            int tx = t % ntx, ty = t / ntx;
            int gx0 = tx * p->xtsiz + p->xtosiz, gy0 = ty * p->ytsiz + p->ytosiz, gx1 = gx0 + p->xtsiz, gy1 = gy0 + p->ytsiz;
            if (gx0 < p->xosiz) gx0 = p->xosiz;
            if (gy0 < p->yosiz) gy0 = p->yosiz;
            if (gx1 > p->xsiz) gx1 = p->xsiz;
            if (gy1 > p->ysiz) gy1 = p->ysiz;
            TileGeom g;
            g.x0 = gx0 - p->xosiz; g.y0 = gy0 - p->yosiz; g.tw = gx1 - gx0; g.th = gy1 - gy0;
            if (g.tw < 0) g.tw = 0;
            if (g.th < 0) g.th = 0;
            g.ox = gx0; g.oy = gy0; g.coeff_off = off;
            off += (long long)g.tw * g.th * p->components;
            out->push_back(g);
        }
    }
    return ntx * nty;
}

int spec_from_fwd(const j2k_fwd_params* p, bool planar, Spec& s) {
    if (!p) return fail(J2K_ERR_INVALID_ARG, "params is NULL");
    int rc = validate_common(p->width, p->height, p->components, p->bit_depth, p->num_levels);
    if (rc) return rc;
    if (p->tile_width < 0 || p->tile_height < 0) return fail(J2K_ERR_INVALID_ARG, "invalid tile size");
    if (p->n_steps < 0 || p->n_steps > J2K_MAX_BANDS) return fail(J2K_ERR_INVALID_ARG, "n_steps out of range");
    s = Spec{};
    s.fwd = true; s.W = p->width; s.H = p->height; s.C = p->components; s.bit_depth = p->bit_depth; s.is_signed = p->is_signed != 0;
    s.L = p->num_levels; s.reversible = p->reversible != 0; s.htj2k = p->htj2k != 0; s.mct_mode = p->mct_mode;
    fwd_tiles(p, &s.tiles);
    for (const TileGeom& g : s.tiles)
        if ((long long)g.tw * g.th * s.C > (1LL << 31) - 1) return fail(J2K_ERR_INVALID_ARG, "tile too large: %dx%d", g.tw, g.th);
    s.n_steps = p->n_steps; memcpy(s.steps, p->steps, sizeof s.steps);
    s.fuse_shift = p->fuse_t1_shift != 0; s.planar_in = planar; s.direct = false; s.want_planes = false;
    s.mct_matrix = p->mct_matrix; s.mct_has_offsets = p->mct_has_offsets; s.mct_offsets = p->mct_offsets;
    s.n_bindings = p->n_bindings; s.bindings = p->bindings;
    return 0;
}

int spec_from_inv(const j2k_inv_params* p, bool want_planes, Spec& s) {
    if (!p) return fail(J2K_ERR_INVALID_ARG, "params is NULL");
    if (p->xtsiz <= 0 || p->ytsiz <= 0) return fail(J2K_ERR_INVALID_ARG, "invalid tile size %dx%d", p->xtsiz, p->ytsiz);
    if (p->xosiz < 0 || p->yosiz < 0 || p->xtosiz < 0 || p->ytosiz < 0 || p->xtosiz > p->xosiz || p->ytosiz > p->yosiz)
        return fail(J2K_ERR_INVALID_ARG, "invalid image / tile offsets");
    int rc = validate_common(p->xsiz - p->xosiz, p->ysiz - p->yosiz, p->components, p->bit_depth, p->num_levels);
    if (rc) return rc;
    if (p->n_steps < 0 || p->n_steps > J2K_MAX_BANDS) return fail(J2K_ERR_INVALID_ARG, "n_steps out of range");
    s = Spec{};
    s.fwd = false; s.W = p->xsiz - p->xosiz; s.H = p->ysiz - p->yosiz; s.C = p->components; s.bit_depth = p->bit_depth;
    s.is_signed = p->is_signed != 0; s.L = p->num_levels; s.reversible = p->reversible != 0; s.htj2k = p->htj2k != 0; s.mct_mode = p->mct_mode;
    inv_tiles(p, &s.tiles);
    for (const TileGeom& g : s.tiles)
        if ((long long)g.tw * g.th * s.C > (1LL << 31) - 1) return fail(J2K_ERR_INVALID_ARG, "tile too large: %dx%d", g.tw, g.th);
    s.n_steps = p->n_steps; memcpy(s.steps, p->steps, sizeof s.steps);
    s.fuse_shift = p->fuse_t1_halve != 0; s.planar_in = false; s.direct = false; s.want_planes = want_planes;
    s.mct_matrix = p->mct_matrix; s.mct_has_offsets = p->mct_has_offsets; s.mct_offsets = p->mct_offsets;
    s.n_bindings = p->n_bindings; s.bindings = p->bindings;
    return 0;
}

// ------------------------------------------------------------------ code-block interface

// Blocks of one tile-component plane in the order buildTilePacketEncoder walks them (encoder.go:2424-2431):
// resolutions 0..L, sub-bands of getSubbandsForResolution (:3059-3197, origin-0 ceiling divisions), blocks of
// partitionIntoCodeBlocks (:3215-3285) row-major over (cby, cbx).
void codeblock_layout(int width, int height, int L, int cbw, int cbh, std::vector<j2k_cblk>& out) {
    auto cdp2 = [](int a, int b) { return (a + (1 << b) - 1) >> b; };
    long long off = 0;
    auto band = [&](int bx0, int by0, int bw, int bh, int bandno, int res) {
        if (bw <= 0 || bh <= 0) return;
        const int ncbx = (bw + cbw - 1) / cbw, ncby = (bh + cbh - 1) / cbh;
        for (int cby = 0; cby < ncby; cby++)
            for (int cbx = 0; cbx < ncbx; cbx++) {
                const int x0 = cbx * cbw, y0 = cby * cbh;
                const int aw = (x0 + cbw > bw ? bw : x0 + cbw) - x0, ah = (y0 + cbh > bh ? bh : y0 + cbh) - y0;
                out.push_back(j2k_cblk{bx0 + x0, by0 + y0, aw, ah, cbx, cby, bandno, res, off});
                off += (long long)aw * ah;
            }
    };
    const int div = 1 << L;
    band(0, 0, (width + div - 1) / div, (height + div - 1) / div, 0, 0);
    for (int res = 1; res <= L; res++) {
        const int level = L - res;
        const int llw = cdp2(width, level + 1), llh = cdp2(height, level + 1);
        band(llw, 0, cdp2(width - (1 << level), level + 1), cdp2(height, level + 1), 1, res);
        band(0, llh, cdp2(width, level + 1), cdp2(height - (1 << level), level + 1), 2, res);
        band(llw, llh, cdp2(width - (1 << level), level + 1), cdp2(height - (1 << level), level + 1), 3, res);
    }
}

int validate_cb(int cbw, int cbh) {  // EncodeParams.Validate, encoder.go:310-316
    auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    if (cbw < 4 || cbw > 1024 || !pow2(cbw)) return fail(J2K_ERR_INVALID_ARG, "invalid code-block width: %d (must be power of 2, 4-1024)", cbw);
    if (cbh < 4 || cbh > 1024 || !pow2(cbh)) return fail(J2K_ERR_INVALID_ARG, "invalid code-block height: %d (must be power of 2, 4-1024)", cbh);
    return 0;
}

struct BlockTable {
    DevBuf tab;
    int nblocks = 0;                 // per frame
    long long coeffs_per_frame = 0;
    ~BlockTable() { tab.release(); }
};

size_t frame_block_count(const std::vector<TileGeom>& tiles, int C, int L, int cbw, int cbh) {
    size_t n = 0;
    std::vector<j2k_cblk> t;
    for (const TileGeom& g : tiles) {
        t.clear();
        codeblock_layout(g.tw, g.th, L, cbw, cbh, t);
        n += t.size() * (size_t)C;
    }
    return n;
}

// Flat per-frame table: every block of every tile-component, with its position in the coefficient planes.
int get_block_table(DeviceCtx& d, const Spec& s, int cbw, int cbh, long long coeffs_per_frame, BlockTable** out) {
    std::string key;
    char buf[96];
    snprintf(buf, sizeof buf, "%d|%d|%d|%d|%lld|", s.C, s.L, cbw, cbh, coeffs_per_frame);
    key = buf;
    for (const TileGeom& g : s.tiles) { snprintf(buf, sizeof buf, "%d,%d,%lld;", g.tw, g.th, g.coeff_off); key += buf; }
    auto it = d.block_tables.find(key);
    if (it != d.block_tables.end()) { *out = it->second.get(); return 0; }
    if (d.block_tables.size() >= 16) d.block_tables.clear();
    std::vector<BlockEntry> host;
    std::vector<j2k_cblk> t;
    for (const TileGeom& g : s.tiles) {
        t.clear();
        codeblock_layout(g.tw, g.th, s.L, cbw, cbh, t);
        for (int c = 0; c < s.C; c++) {
            const long long plane = g.coeff_off + (long long)c * g.tw * g.th;
            for (const j2k_cblk& b : t) {
                BlockEntry e;
                e.plane_off = plane + (long long)b.y0 * g.tw + b.x0;
                e.block_off = plane + b.offset;
                e.stride = g.tw; e.w = b.width; e.h = b.height;
                e.vec = (b.width % 4 == 0 && g.tw % 4 == 0 && e.plane_off % 4 == 0 && e.block_off % 4 == 0 && coeffs_per_frame % 4 == 0) ? 1 : 0;
                e.comp = c; e.pad_ = 0;
                host.push_back(e);
            }
        }
    }
    std::unique_ptr<BlockTable> T(new BlockTable());
    T->nblocks = (int)host.size();
    T->coeffs_per_frame = coeffs_per_frame;
    int rc = T->tab.ensure(host.size() * sizeof(BlockEntry) + 16);
    if (rc) return rc;
    if (!host.empty()) CK(cudaMemcpy(T->tab.p, host.data(), host.size() * sizeof(BlockEntry), cudaMemcpyHostToDevice));
    CK(cudaDeviceSynchronize());  // the table is read from non-blocking streams (see get_plan)
    *out = T.get();
    d.block_tables[key] = std::move(T);
    return 0;
}

int launch_gather(j2k_ctx* ctx, BlockTable& T, int nframes, const int32_t* d_coeffs, int32_t* d_blocks, int32_t* d_numbps, bool sub6,
                  cudaStream_t st) {
    const long long total = (long long)T.nblocks * nframes;
    if (total <= 0) return 0;
    const unsigned grid = (unsigned)((total + 3) / 4);
    // the 128-bit row copies need 16-byte aligned planes (a sliced device tensor may not be): scalar copies otherwise
    const int vec_ok = ((((uintptr_t)d_coeffs) | ((uintptr_t)d_blocks)) & 15u) == 0;
    J2K_LAUNCH(gather_blocks_kernel, grid, 128, st, (const int*)d_coeffs, T.coeffs_per_frame, (const BlockEntry*)T.tab.p, T.nblocks, total,
               (int*)d_blocks, (int*)d_numbps, sub6 ? 1 : 0, vec_ok);
    CK(cudaGetLastError());
    ctx->launches++;
    return 0;
}

// HT code-blocks hold at most 4096 samples (ISO/IEC 15444-15 restricts xcb + ycb <= 12 as Part 1 does)
int validate_ht_cb(int cbw, int cbh) {
    int rc = validate_cb(cbw, cbh);
    if (rc) return rc;
    if ((long long)cbw * cbh > 4096) return fail(J2K_ERR_UNSUPPORTED, "HT code-blocks hold at most 4096 samples (%d x %d)", cbw, cbh);
    return 0;
}

// Two launches (j2k_ht.cuh): VLC / MEL / UVLC with one thread per code-block into the device's scratch (4 B per quad), then
// MagSgn with one warp per code-block; to_planes: decoded samples go straight into the Mallat planes (assembleSubbands).
// The scratch is one buffer per device: a launch on another stream waits for the previous one's MagSgn kernel.
int launch_ht_decode(j2k_ctx* ctx, DeviceCtx& d, BlockTable& T, int cbw, int cbh, int nframes, const unsigned char* d_bytes,
                     const HtBlock* d_descs, int32_t* d_out, int to_planes, int32_t* d_status, cudaStream_t st) {
    const long long total = (long long)T.nblocks * nframes;
    if (total <= 0) return 0;
    const int scw = ht_scratch_words(cbw, cbh), qs = (cbw + 1) / 2;
    int rc = d.ht_scratch.ensure((size_t)total * scw * 4 + 16);
    if (rc) return rc;
    if (!d.ev_ht) CK(cudaEventCreateWithFlags(&d.ev_ht, cudaEventDisableTiming));
    if (d.ev_ht_used) CK(cudaStreamWaitEvent(st, d.ev_ht, 0));
    J2K_LAUNCH(ht_vlc_kernel, (unsigned)((total + 31) / 32), 32, st, d_bytes, d_descs, (const BlockEntry*)T.tab.p, T.nblocks, total,
               (unsigned*)d.ht_scratch.p, scw, qs);
    CK(cudaGetLastError());
    const int warps = 4;
    const int wsm = ht_warp_smem(cbw);
    J2K_LAUNCH_SMEM(ht_magsgn_kernel, (unsigned)((total + warps - 1) / warps), warps * 32, warps * wsm, st, d_bytes, d_descs,
                    (const BlockEntry*)T.tab.p, T.nblocks, total, T.coeffs_per_frame, (const unsigned*)d.ht_scratch.p, scw, qs, (int*)d_out,
                    to_planes, (int*)d_status, wsm);
    CK(cudaGetLastError());
    CK(cudaEventRecord(d.ev_ht, st));
    d.ev_ht_used = true;
    ctx->launches += 2;
    return 0;
}

// Per-block Kmax in block-table order from the caller's [components][3 * levels + 1] table (bandNumbps per sub-band:
// index 0 = LL, then HL, LH, HH from the coarsest resolution; Encoder.bandNumbps, encoder.go:3303).
int ht_block_kmax(const Spec& s, int cbw, int cbh, const uint8_t* kmax, std::vector<unsigned char>& out, int& kmax_max) {
    out.clear();
    kmax_max = 0;
    const int nb = 3 * s.L + 1;
    std::vector<j2k_cblk> t;
    for (const TileGeom& g : s.tiles) {
        t.clear();
        codeblock_layout(g.tw, g.th, s.L, cbw, cbh, t);
        for (int c = 0; c < s.C; c++)
            for (const j2k_cblk& b : t) {
                const int idx = b.res == 0 ? 0 : 1 + 3 * (b.res - 1) + (b.band - 1);
                const int k = kmax[c * nb + idx];
                // "invalid HTJ2K Kmax" (openjph_cleanup_encoder.go:201-203) / "requires Kmax coding context" (encoder.go:64-66)
                if (k <= 0 || k >= 31) return fail(J2K_ERR_INVALID_ARG, "invalid HTJ2K Kmax: %d (component %d, sub-band %d)", k, c, idx);
                if (k > kmax_max) kmax_max = k;
                out.push_back((unsigned char)k);
            }
    }
    return 0;
}

// Four launches (j2k_ht_enc.cuh).  d_kmax: per-block Kmax (block-table order) on the device; d_offsets: total + 1 words, the
// last one receives the size of the compact stream; nothing is written to d_bytes beyond `cap`.
int launch_ht_encode(j2k_ctx* ctx, DeviceCtx& d, BlockTable& T, int cbw, int cbh, int kmax_max, int nframes, const int32_t* d_coeffs,
                     const unsigned char* d_kmax, unsigned char* d_bytes, unsigned long long cap, HtBlock* d_recs,
                     unsigned long long* d_offsets, cudaStream_t st) {
    const long long total = (long long)T.nblocks * nframes;
    if (total <= 0) return 0;
    const HtEncLayout L = ht_enc_layout(cbw, cbh, kmax_max);
    int rc = d.he_slots.ensure((size_t)total * L.slot_bytes + 64);
    if (rc) return rc;
    if ((rc = d.he_info.ensure((size_t)total * sizeof(HtEncInfo) + 64))) return rc;
    const int warps = 4, wsm = ht_enc_warp_smem(cbw);
    J2K_LAUNCH_SMEM(ht_enc_quads_kernel, (unsigned)((total + warps - 1) / warps), warps * 32, warps * wsm, st, (const int*)d_coeffs,
                    T.coeffs_per_frame, (const BlockEntry*)T.tab.p, d_kmax, T.nblocks, total, (unsigned char*)d.he_slots.p, L,
                    (HtEncInfo*)d.he_info.p, wsm);
    CK(cudaGetLastError());
    J2K_LAUNCH(ht_enc_pack_kernel, (unsigned)((total + 31) / 32), 32, st, total, (unsigned char*)d.he_slots.p, L, (HtEncInfo*)d.he_info.p);
    CK(cudaGetLastError());
    J2K_LAUNCH(ht_enc_scan_kernel, 1, 1024, st, (const HtEncInfo*)d.he_info.p, total, d_offsets);
    CK(cudaGetLastError());
    J2K_LAUNCH(ht_enc_compact_kernel, (unsigned)((total + warps - 1) / warps), warps * 32, st, total, (const unsigned char*)d.he_slots.p, L,
               (const HtEncInfo*)d.he_info.p, (const unsigned long long*)d_offsets, d_kmax, T.nblocks, d_bytes, cap, d_recs);
    CK(cudaGetLastError());
    ctx->launches += 4;
    return 0;
}

// upper bound of the compact stream of `total` blocks (what the per-block output areas can hold)
size_t ht_enc_bound(int cbw, int cbh, int kmax_max, long long total) {
    const HtEncLayout L = ht_enc_layout(cbw, cbh, kmax_max);
    return (size_t)total * (size_t)(L.slot_bytes - L.ms_out_off);
}

// Per-component MaxShift values of the caller (NULL: no ROI) -> kernel argument.
int make_roi(const int32_t* roi_maxshift, int C, RoiShifts& r) {
    memset(&r, 0, sizeof r);
    if (!roi_maxshift) return 0;
    if (C > J2K_ROI_MAXC) return fail(J2K_ERR_UNSUPPORTED, "ROI shifts for %d components (at most %d)", C, J2K_ROI_MAXC);
    for (int c = 0; c < C; c++) {
        if (roi_maxshift[c] < 0 || roi_maxshift[c] > 255) return fail(J2K_ERR_INVALID_ARG, "invalid ROI shift: %d (must be <=255)", roi_maxshift[c]);
        r.shift[c] = roi_maxshift[c];
        if (roi_maxshift[c]) r.any = 1;
    }
    return 0;
}

int launch_scatter(j2k_ctx* ctx, BlockTable& T, int nframes, const int32_t* d_blocks, int32_t* d_coeffs, cudaStream_t st,
                   const RoiShifts& roi) {
    const long long total = (long long)T.nblocks * nframes;
    if (total <= 0) return 0;
    const unsigned grid = (unsigned)((total + 3) / 4);
    const int vec_ok = ((((uintptr_t)d_coeffs) | ((uintptr_t)d_blocks)) & 15u) == 0;
    J2K_LAUNCH(scatter_blocks_kernel, grid, 128, st, (const int*)d_blocks, T.coeffs_per_frame, (const BlockEntry*)T.tab.p, T.nblocks, total,
               (int*)d_coeffs, roi, vec_ok);
    CK(cudaGetLastError());
    ctx->launches++;
    return 0;
}

// A plan owns its scratch (LL ping-pong planes, job-control block): plans are cached per stream so that launches the
// caller enqueues on different streams may overlap (the tail of one batch under the head of the next).
// Cache key of a parameter block: only what the plan depends on.  Callers (C, Go) hand over structs whose padding, reserved
// words and unused tail entries (steps beyond n_steps, bindings beyond n_bindings, the matrix of an unused MCT mode) hold
// whatever was on their stack; keyed on the raw bytes every such call would build a new plan.
template <typename P>
std::string canonical_params(const P& p, bool is_fwd) {
    P k;
    memset(&k, 0, sizeof k);
    if constexpr (std::is_same<P, j2k_fwd_params>::value) {
        k.width = p.width; k.height = p.height; k.tile_width = p.tile_width; k.tile_height = p.tile_height;
        k.fuse_t1_shift = p.reversible && !p.htj2k ? (p.fuse_t1_shift != 0) : 0;
    } else {
        k.xsiz = p.xsiz; k.ysiz = p.ysiz; k.xosiz = p.xosiz; k.yosiz = p.yosiz;
        k.xtsiz = p.xtsiz; k.ytsiz = p.ytsiz; k.xtosiz = p.xtosiz; k.ytosiz = p.ytosiz;
        k.fuse_t1_halve = p.reversible && !p.htj2k ? (p.fuse_t1_halve != 0) : 0;
    }
    (void)is_fwd;
    k.components = p.components; k.bit_depth = p.bit_depth; k.is_signed = p.is_signed != 0; k.num_levels = p.num_levels;
    k.reversible = p.reversible != 0; k.htj2k = p.htj2k != 0; k.mct_mode = p.mct_mode;
    const int C = p.components < 0 ? 0 : (p.components > J2K_MAX_COMPONENTS ? J2K_MAX_COMPONENTS : p.components);
    if (p.mct_mode == J2K_MCT_CUSTOM_INT || p.mct_mode == J2K_MCT_CUSTOM_Q13 || p.mct_mode == J2K_MCT_CUSTOM_FLOAT) {
        for (int i = 0; i < C * C; i++) k.mct_matrix[i] = p.mct_matrix[i];
        k.mct_has_offsets = p.mct_has_offsets != 0;
        if (k.mct_has_offsets) for (int i = 0; i < C; i++) k.mct_offsets[i] = p.mct_offsets[i];
    }
    if (p.mct_mode == J2K_MCT_BINDINGS) {
        k.n_bindings = p.n_bindings;
        for (int b = 0; b < p.n_bindings && b < J2K_MAX_BINDINGS; b++) {
            const j2k_mct_binding& q = p.bindings[b];
            j2k_mct_binding& o = k.bindings[b];
            o.n_components = q.n_components; o.element_type = q.element_type; o.has_matrix = q.has_matrix != 0; o.has_offsets = q.has_offsets != 0;
            const int n = q.n_components > 0 && q.n_components <= J2K_MAX_COMPONENTS ? q.n_components : C;
            for (int i = 0; i < q.n_components && i < J2K_MAX_COMPONENTS; i++) o.component_ids[i] = q.component_ids[i];
            if (o.has_matrix) for (int i = 0; i < n * n; i++) o.matrix[i] = q.matrix[i];
            if (o.has_offsets) for (int i = 0; i < n; i++) o.offsets[i] = q.offsets[i];
        }
    }
    if (!p.reversible) {
        k.n_steps = p.n_steps;
        for (int i = 0; i < p.n_steps && i < J2K_MAX_BANDS; i++) k.steps[i] = p.steps[i];
    }
    return std::string((const char*)&k, sizeof k);
}

int get_plan(DeviceCtx& d, const Spec& s, const void* pblob, size_t pbytes, int nframes, long long frame_samples, Plan** out,
             const void* stream_tag = nullptr) {
    std::string key;
    if (pbytes == sizeof(j2k_fwd_params) && s.fwd && !s.direct) key = canonical_params(*(const j2k_fwd_params*)pblob, true);
    else if (pbytes == sizeof(j2k_inv_params) && !s.fwd && !s.direct) key = canonical_params(*(const j2k_inv_params*)pblob, false);
    else key.assign((const char*)pblob, pbytes);
    char tail[128];
    snprintf(tail, sizeof tail, "|%d|%d|%lld|%d|%d|%d|%p", s.fwd ? 1 : 0, nframes, frame_samples, s.planar_in ? 1 : 0, s.want_planes ? 1 : 0,
             s.direct ? 1 : 0, stream_tag);
    key += tail;
    auto it = d.plans.find(key);
    if (it != d.plans.end()) { it->second->last_use = ++d.use_clock; *out = it->second.get(); return 0; }
    if (d.plans.size() >= 24) {  // evict the least recently used plan (its buffers are freed: cudaFree waits for the device)
        auto old = d.plans.begin();
        for (auto jt = d.plans.begin(); jt != d.plans.end(); ++jt)
            if (jt->second->last_use < old->second->last_use) old = jt;
        d.plans.erase(old);
    }
    std::unique_ptr<Plan> P(new Plan());
    int rc = build_plan(s, nframes, frame_samples, *P);
    if (rc) return rc;
    // build_plan uploads its tables and zeroes the job-control block with plain cudaMemcpy / cudaMemset (legacy stream, pageable
    // sources); the plan runs on non-blocking streams that are not ordered against those: finish them before the first launch
    CK(cudaDeviceSynchronize());
    P->last_use = ++d.use_clock;
    *out = P.get();
    d.plans[key] = std::move(P);
    return 0;
}

int set_dev(j2k_ctx* ctx, int di) {
    if (!ctx) return fail(J2K_ERR_INVALID_ARG, "context is NULL");
    if (di < 0 || di >= (int)ctx->devs.size()) return fail(J2K_ERR_INVALID_ARG, "device index %d out of range", di);
    CK(cudaSetDevice(ctx->devs[di].dev));
    return 0;
}

// ------------------------------------------------------------------ pageable callers: pinned staging

// Is `p` page-locked host memory (j2k_acquire_buffer, cudaHostAlloc, cudaHostRegister)?  Anything else - a Go slice, a
// numpy array, malloc - is pageable and goes through the staging ring.
bool host_pinned(const void* p) {
#ifdef J2K_EMU
    (void)p;
    return getenv("J2K_EMU_PAGEABLE") == nullptr;
#else
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
#endif
}

void pool_start(j2k_ctx* ctx) {
    if (!ctx->workers.empty()) return;
    unsigned hw = std::thread::hardware_concurrency();
    int n = env_int("J2K_STAGE_THREADS", hw >= 32 ? 8 : (hw >= 8 ? 4 : 2));
    if (n < 1) n = 1;
    for (int i = 0; i < n; i++)
        ctx->workers.emplace_back([ctx] {
            for (;;) {
                std::function<void()> f;
                {
                    std::unique_lock<std::mutex> lk(ctx->wq_mu);
                    ctx->wq_cv.wait(lk, [&] { return ctx->wq_stop || !ctx->wq.empty(); });
                    if (ctx->wq.empty()) return;  // stop requested and nothing left
                    f = std::move(ctx->wq.front());
                    ctx->wq.pop_front();
                }
                f();
                {
                    std::lock_guard<std::mutex> lk(ctx->wq_mu);
                    ctx->wq_inflight--;
                }
                ctx->wq_done.notify_all();
            }
        });
}

void pool_stop(j2k_ctx* ctx) {
    {
        std::lock_guard<std::mutex> lk(ctx->wq_mu);
        ctx->wq_stop = true;
    }
    ctx->wq_cv.notify_all();
    for (auto& t : ctx->workers) t.join();
    ctx->workers.clear();
}

// memcpy split over the staging threads (the caller waits: the chunk is about to be handed to the copy engine / the user)
void pool_copy(j2k_ctx* ctx, void* dst, const void* src, size_t n) {
    const size_t nw = ctx->workers.size();
    if (nw <= 1 || n < (1u << 20)) { memcpy(dst, src, n); return; }
    size_t piece = ((n + nw - 1) / nw + 4095) & ~(size_t)4095;
    {
        std::lock_guard<std::mutex> lk(ctx->wq_mu);
        for (size_t off = 0; off < n; off += piece) {
            const size_t m = off + piece <= n ? piece : n - off;
            unsigned char* d = (unsigned char*)dst + off;
            const unsigned char* sp = (const unsigned char*)src + off;
            ctx->wq.push_back([d, sp, m] { memcpy(d, sp, m); });
            ctx->wq_inflight++;
        }
    }
    ctx->wq_cv.notify_all();
    std::unique_lock<std::mutex> lk(ctx->wq_mu);
    ctx->wq_done.wait(lk, [&] { return ctx->wq_inflight == 0; });
}

int stager_init(j2k_ctx* ctx, DeviceCtx& d) {
    Stager& S = d.stg;
    if (S.ready) return 0;
    pool_start(ctx);
    S.CHUNK = (size_t)env_int("J2K_STAGE_CHUNK_KB", 4096) << 10;
    if (S.CHUNK < 1024) S.CHUNK = 1024;
    for (int k = 0; k < Stager::NSLOT; k++) {
        CK(cudaHostAlloc((void**)&S.up[k], S.CHUNK, cudaHostAllocPortable));
        CK(cudaHostAlloc((void**)&S.down[k], S.CHUNK, cudaHostAllocPortable));
        CK(cudaEventCreateWithFlags(&S.up_ev[k], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&S.down_ev[k], cudaEventDisableTiming));
    }
    S.ready = true;
    return 0;
}

void stager_release(DeviceCtx& d) {
    Stager& S = d.stg;
    for (int k = 0; k < Stager::NSLOT; k++) {
        if (S.up[k]) cudaFreeHost(S.up[k]);
        if (S.down[k]) cudaFreeHost(S.down[k]);
        if (S.up_ev[k]) cudaEventDestroy(S.up_ev[k]);
        if (S.down_ev[k]) cudaEventDestroy(S.down_ev[k]);
        S.up[k] = S.down[k] = nullptr; S.up_ev[k] = S.down_ev[k] = nullptr;
    }
    S.ready = false;
}

// pageable host -> device on the upload stream, chunk by chunk through the pinned ring
int stage_up(j2k_ctx* ctx, DeviceCtx& d, void* dst, const unsigned char* src, size_t bytes) {
    int rc = stager_init(ctx, d);
    if (rc) return rc;
    Stager& S = d.stg;
    for (size_t off = 0; off < bytes; off += S.CHUNK) {
        const size_t n = off + S.CHUNK <= bytes ? S.CHUNK : bytes - off;
        const int k = S.up_next;
        S.up_next = (k + 1) % Stager::NSLOT;
        if (S.up_busy[k]) CK(cudaEventSynchronize(S.up_ev[k]));  // the copy engine has drained this chunk
        pool_copy(ctx, S.up[k], src + off, n);
        CK(cudaMemcpyAsync((unsigned char*)dst + off, S.up[k], n, cudaMemcpyHostToDevice, d.s_h2d));
        CK(cudaEventRecord(S.up_ev[k], d.s_h2d));
        S.up_busy[k] = true;
    }
    return 0;
}

// device -> pageable host for everything queued in S.pending: a sliding window of NSLOT chunks in flight on the download
// stream, each copied out to the caller's buffer as soon as its event has fired
int stage_drain(j2k_ctx* ctx, DeviceCtx& d) {
    Stager& S = d.stg;
    if (S.pending.empty()) return 0;
    int rc = stager_init(ctx, d);
    if (rc) return rc;
    struct Chunk { unsigned char* dst; const unsigned char* src; size_t n; };
    std::vector<Chunk> ch;
    for (const Stager::Pending& p : S.pending)
        for (size_t off = 0; off < p.bytes; off += S.CHUNK)
            ch.push_back({p.dst + off, p.src + off, off + S.CHUNK <= p.bytes ? S.CHUNK : p.bytes - off});
    S.pending.clear();
    size_t issued = 0, done = 0;
    while (done < ch.size()) {
        while (issued < ch.size() && issued - done < (size_t)Stager::NSLOT) {
            const int k = (int)(issued % Stager::NSLOT);
            CK(cudaMemcpyAsync(S.down[k], ch[issued].src, ch[issued].n, cudaMemcpyDeviceToHost, d.s_d2h));
            CK(cudaEventRecord(S.down_ev[k], d.s_d2h));
            issued++;
        }
        const int k = (int)(done % Stager::NSLOT);
        CK(cudaEventSynchronize(S.down_ev[k]));
        pool_copy(ctx, ch[done].dst, S.down[k], ch[done].n);
        done++;
    }
    return 0;
}

// Frames [f0, f1) of a host batch on device di: H2D, run, D2H, double-buffered in sub-batches.
struct HostJob {
    bool fwd; bool planar;
    const Spec* spec; const void* pblob; size_t pbytes;
    const unsigned char* h_pix_in; unsigned char* h_pix_out; size_t frame_stride_bytes;
    const int32_t* h_coef_in; int32_t* h_coef_out;
    int32_t* h_planes;
    size_t pix_bytes_per_frame; long long coeffs_per_frame;
    int cb_w = 0, cb_h = 0;      // > 0: coefficients cross the boundary block-major (code-block interface)
    int32_t* h_numbps = nullptr; // forward, block mode: cblkNumbps per block
    bool may_stage = true;       // synchronous call: pageable buffers go through the pinned staging ring (async: rejected)
    RoiShifts roi{};             // inverse, block mode: per-component MaxShift applied while scattering
    const int32_t* h_block_shift = nullptr;      // inverse, block mode: general-scaling shift per (frame, block), or NULL
    const unsigned char* h_sample_mask = nullptr; // and its optional per-sample mask (block-major, one byte per coefficient)
    // inverse, block mode, HTJ2K: the blocks arrive as HT cleanup segments and are decoded on the device (j2k_inverse_ht)
    const unsigned char* h_ht_bytes = nullptr; const j2k_ht_cblk* h_ht_recs = nullptr; int32_t* h_ht_status = nullptr;
};

int enqueue_host_job(j2k_ctx* ctx, int di, const HostJob& J, int f0, int f1, bool timing) {
    DeviceCtx& d = ctx->devs[di];
    int rc = set_dev(ctx, di);
    if (rc) return rc;
    const int n = f1 - f0;
    if (n <= 0) return 0;
    const Spec& s = *J.spec;
    const long long HW = (long long)s.W * s.H;
    // sub-batches: enough frames per launch to fill the GPU, at least two to overlap copies with kernels
    long long frame_samples_total = HW * s.C;
    // ~16 Msamples per sub-batch: the first upload (the only copy nothing hides) stays short, a launch still fills the GPU
    static const long long sub_samples = env_int("J2K_SUBBATCH_SAMPLES", 0) > 0 ? (long long)env_int("J2K_SUBBATCH_SAMPLES", 0)
                                                                                   : (long long)env_int("J2K_SUBBATCH_MSAMPLES", 16) << 20;
    int sub = (int)(sub_samples / (frame_samples_total > 0 ? frame_samples_total : 1));
    if (sub < 1) sub = 1;
    if (sub > n) sub = n;
    if (timing) CK(cudaEventRecord(d.ev_t[0], d.s_h2d));
    // caller-owned pageable memory (a Go slice, a numpy array) moves through the pinned staging ring; pinned buffers
    // (j2k_acquire_buffer) are handed to the copy engine as they are
    const bool page_in = !host_pinned(J.fwd ? (const void*)J.h_pix_in : J.h_ht_recs ? (const void*)J.h_ht_bytes : (const void*)J.h_coef_in);
    const bool page_out = !host_pinned(J.fwd ? (const void*)J.h_coef_out : (const void*)J.h_pix_out) ||
                          (J.h_planes && !host_pinned(J.h_planes)) || (J.h_numbps && !host_pinned(J.h_numbps)) ||
                          (J.h_ht_status && !host_pinned(J.h_ht_status));
    if ((page_in || page_out) && !J.may_stage)
        return fail(J2K_ERR_INVALID_ARG, "asynchronous calls need buffers from j2k_acquire_buffer (pinned); this one is pageable");
    auto up = [&](void* dst, const unsigned char* src, size_t bytes) -> int {
        if (page_in) return stage_up(ctx, d, dst, src, bytes);
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, d.s_h2d));
        return 0;
    };
    auto down = [&](void* dst, const void* src, size_t bytes) -> int {
        if (page_out) { d.stg.pending.push_back({(unsigned char*)dst, (const unsigned char*)src, bytes}); return 0; }
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, d.s_d2h));
        return 0;
    };
    int it = 0;
    for (int b = f0; b < f1; b += sub, it++) {
        const int nb = (b + sub <= f1) ? sub : f1 - b;
        const int slot = it & 1;
        const size_t in_bytes = J.fwd ? (J.planar ? (size_t)nb * s.C * HW * 4 : (size_t)nb * J.pix_bytes_per_frame) : (size_t)nb * J.coeffs_per_frame * 4;
        const size_t out_bytes = J.fwd ? (size_t)nb * J.coeffs_per_frame * 4 : (size_t)nb * J.pix_bytes_per_frame;
        if ((rc = d.in[slot].ensure(in_bytes))) return rc;
        if ((rc = d.out[slot].ensure(out_bytes))) return rc;
        if (J.h_planes && (rc = d.planes[slot].ensure((size_t)nb * s.C * HW * 4))) return rc;
        BlockTable* BT = nullptr;
        if (J.cb_w > 0) {
            if ((rc = get_block_table(d, s, J.cb_w, J.cb_h, J.coeffs_per_frame, &BT))) return rc;
            if (!J.h_ht_recs && (rc = d.blk[slot].ensure((size_t)nb * J.coeffs_per_frame * 4))) return rc;
            if (J.fwd && (rc = d.nbp[slot].ensure((size_t)nb * BT->nblocks * 4 + 16))) return rc;
        }
        // the slot's previous kernel must have consumed `in`, its previous D2H must have drained `out`
        if (d.ev_k_used[slot]) CK(cudaStreamWaitEvent(d.s_h2d, d.ev_k[slot], 0));
        if (J.fwd) {
            if (J.planar || J.frame_stride_bytes == J.pix_bytes_per_frame) {
                const unsigned char* src = J.planar ? J.h_pix_in : J.h_pix_in + (size_t)b * J.frame_stride_bytes;
                if ((rc = up(d.in[slot].p, src, in_bytes))) return rc;
            } else if (page_in) {
                for (int f = 0; f < nb; f++)
                    if ((rc = up((unsigned char*)d.in[slot].p + (size_t)f * J.pix_bytes_per_frame,
                                 J.h_pix_in + (size_t)(b + f) * J.frame_stride_bytes, J.pix_bytes_per_frame))) return rc;
            } else {
                CK(cudaMemcpy2DAsync(d.in[slot].p, J.pix_bytes_per_frame, J.h_pix_in + (size_t)b * J.frame_stride_bytes, J.frame_stride_bytes,
                                     J.pix_bytes_per_frame, nb, cudaMemcpyHostToDevice, d.s_h2d));
            }
        } else if (J.h_ht_recs) {
            // this sub-batch's cleanup segments (the byte range its records span) and the records, re-based to that range
            const size_t count = (size_t)nb * BT->nblocks;
            const j2k_ht_cblk* src = J.h_ht_recs + (size_t)b * BT->nblocks;
            unsigned long long lo = ~0ull, hi = 0;
            for (size_t i = 0; i < count; i++)
                if (src[i].length) {
                    if (src[i].offset < lo) lo = src[i].offset;
                    if (src[i].offset + src[i].length > hi) hi = src[i].offset + src[i].length;
                }
            if (hi == 0) lo = 0;
            std::vector<HtBlock> r(count);
            for (size_t i = 0; i < count; i++) {
                r[i].offset = src[i].length ? src[i].offset - lo : 0;
                r[i].length = src[i].length; r[i].kmax = src[i].kmax; r[i].mmsb = src[i].missing_msbs; r[i].reserved = 0;
            }
            if ((rc = d.ht_bytes[slot].ensure((size_t)(hi - lo) + 16))) return rc;
            if ((rc = d.ht_desc[slot].ensure(count * sizeof(HtBlock) + 16))) return rc;
            if (J.h_ht_status && (rc = d.ht_status[slot].ensure(count * 4 + 16))) return rc;
            if (hi > lo && (rc = up(d.ht_bytes[slot].p, J.h_ht_bytes + lo, (size_t)(hi - lo)))) return rc;
            // the re-based records live in `r`: always through the pinned staging ring, which copies them before it returns
            if ((rc = stage_up(ctx, d, d.ht_desc[slot].p, (const unsigned char*)r.data(), count * sizeof(HtBlock)))) return rc;
        } else {
            if ((rc = up(BT ? d.blk[slot].p : d.in[slot].p, (const unsigned char*)(J.h_coef_in + (size_t)b * J.coeffs_per_frame), in_bytes))) return rc;
        }
        CK(cudaEventRecord(d.ev_in[slot], d.s_h2d));
        CK(cudaStreamWaitEvent(d.s_main, d.ev_in[slot], 0));
        if (d.ev_out_used[slot]) CK(cudaStreamWaitEvent(d.s_main, d.ev_out[slot], 0));
        Plan* P = nullptr;
        long long fs = J.fwd ? (J.planar ? 0 : (long long)(J.pix_bytes_per_frame / (s.bit_depth <= 8 ? 1 : 2))) : (long long)(J.pix_bytes_per_frame / (s.bit_depth <= 8 ? 1 : 2));
        if ((rc = get_plan(d, s, J.pblob, J.pbytes, nb, fs, &P))) return rc;
        if (timing && it == 0) CK(cudaEventRecord(d.ev_t[1], d.s_main));
        if (!J.fwd && BT && J.h_ht_recs) {
            // HT block decoding straight into the coefficient planes (assembleSubbands fused into the decoder's stores)
            if ((rc = launch_ht_decode(ctx, d, *BT, J.cb_w, J.cb_h, nb, (const unsigned char*)d.ht_bytes[slot].p, (const HtBlock*)d.ht_desc[slot].p,
                                       (int32_t*)d.in[slot].p, 1, J.h_ht_status ? (int32_t*)d.ht_status[slot].p : nullptr, d.s_main))) return rc;
        } else if (!J.fwd && BT) {
            RoiShifts roi = J.roi;
            if (J.h_block_shift) {  // general scaling: this sub-batch's per-block shifts (and mask) go up behind the blocks
                const size_t nsh = (size_t)nb * BT->nblocks * 4;
                if ((rc = d.gsh[slot].ensure(nsh + 16))) return rc;
                if ((rc = up(d.gsh[slot].p, (const unsigned char*)(J.h_block_shift + (size_t)b * BT->nblocks), nsh))) return rc;
                roi.block_shift = (const int*)d.gsh[slot].p;
                if (J.h_sample_mask) {
                    const size_t nmk = (size_t)nb * J.coeffs_per_frame;
                    if ((rc = d.gmk[slot].ensure(nmk + 16))) return rc;
                    if ((rc = up(d.gmk[slot].p, J.h_sample_mask + (size_t)b * J.coeffs_per_frame, nmk))) return rc;
                    roi.sample_mask = (const unsigned char*)d.gmk[slot].p;
                }
                CK(cudaEventRecord(d.ev_in[slot], d.s_h2d));
                CK(cudaStreamWaitEvent(d.s_main, d.ev_in[slot], 0));
            }
            if ((rc = launch_scatter(ctx, *BT, nb, (const int32_t*)d.blk[slot].p, (int32_t*)d.in[slot].p, d.s_main, roi))) return rc;
        }
        if (J.fwd) rc = run_plan(ctx, *P, d.in[slot].p, d.out[slot].p, nullptr, J.planar, d.s_main);
        else rc = run_plan(ctx, *P, d.out[slot].p, d.in[slot].p, J.h_planes ? d.planes[slot].p : nullptr, false, d.s_main);
        if (rc < 0) return rc;
        if (J.fwd && BT && (rc = launch_gather(ctx, *BT, nb, (const int32_t*)d.out[slot].p, (int32_t*)d.blk[slot].p, (int32_t*)d.nbp[slot].p,
                                               !s.htj2k, d.s_main))) return rc;
        CK(cudaEventRecord(d.ev_k[slot], d.s_main));
        d.ev_k_used[slot] = true;
        // pageable output: the previous sub-batch's results leave now, while this sub-batch computes (its own slot is free
        // again before the sub-batch after this one is enqueued)
        if (page_out && (rc = stage_drain(ctx, d))) return rc;
        CK(cudaStreamWaitEvent(d.s_d2h, d.ev_k[slot], 0));
        if (J.fwd) {
            if ((rc = down(J.h_coef_out + (size_t)b * J.coeffs_per_frame, BT ? d.blk[slot].p : d.out[slot].p, out_bytes))) return rc;
            if (BT && J.h_numbps && (rc = down(J.h_numbps + (size_t)b * BT->nblocks, d.nbp[slot].p, (size_t)nb * BT->nblocks * 4))) return rc;
        } else {
            if (J.frame_stride_bytes == J.pix_bytes_per_frame) {
                if ((rc = down(J.h_pix_out + (size_t)b * J.frame_stride_bytes, d.out[slot].p, out_bytes))) return rc;
            } else if (page_out) {
                for (int f = 0; f < nb; f++)
                    if ((rc = down(J.h_pix_out + (size_t)(b + f) * J.frame_stride_bytes, (unsigned char*)d.out[slot].p + (size_t)f * J.pix_bytes_per_frame,
                                   J.pix_bytes_per_frame))) return rc;
            } else {
                CK(cudaMemcpy2DAsync(J.h_pix_out + (size_t)b * J.frame_stride_bytes, J.frame_stride_bytes, d.out[slot].p, J.pix_bytes_per_frame,
                                     J.pix_bytes_per_frame, nb, cudaMemcpyDeviceToHost, d.s_d2h));
            }
            if (J.h_planes && (rc = down(J.h_planes + (size_t)b * s.C * HW, d.planes[slot].p, (size_t)nb * s.C * HW * 4))) return rc;
            if (J.h_ht_status && (rc = down(J.h_ht_status + (size_t)b * BT->nblocks, d.ht_status[slot].p, (size_t)nb * BT->nblocks * 4))) return rc;
        }
        if (!page_out) {
            CK(cudaEventRecord(d.ev_out[slot], d.s_d2h));
            d.ev_out_used[slot] = true;
        }
    }
    if (page_out && (rc = stage_drain(ctx, d))) return rc;
    if (timing) {
        CK(cudaEventRecord(d.ev_t[2], d.s_main));
        CK(cudaEventRecord(d.ev_t[3], d.s_d2h));
    }
    return 0;
}

int sync_dev(j2k_ctx* ctx, int di) {
    int rc = set_dev(ctx, di);
    if (rc) return rc;
    DeviceCtx& d = ctx->devs[di];
    CK(cudaStreamSynchronize(d.s_h2d));
    CK(cudaStreamSynchronize(d.s_main));
    CK(cudaStreamSynchronize(d.s_d2h));
    return 0;
}

// Failure detection: after a failing CUDA call on device slot `di`, is the device itself gone?  A sticky error
// (illegal address, ECC, a GPU that fell off the bus) makes every later call on it fail, cudaDeviceSynchronize included; a
// logic error (bad launch configuration, a missing kernel variant) leaves the device usable.  A lost device is marked and
// stays out of the round-robin for the life of the context.  J2K_FAULT_DEVICE=<slot> injects such a failure (tests).
int fault_slot() {
    const char* v = getenv("J2K_FAULT_DEVICE");
    return v && *v ? atoi(v) : -1;
}
bool device_lost(j2k_ctx* ctx, int di) {
    DeviceCtx& d = ctx->devs[di];
    if (d.failed) return true;
    bool lost = fault_slot() == di;
    if (!lost) {
        if (cudaSetDevice(d.dev) != cudaSuccess) lost = true;
        else if (cudaDeviceSynchronize() != cudaSuccess) lost = true;
        cudaGetLastError();
    }
    if (lost) {
        d.failed = true;
        char msg[160];
        snprintf(msg, sizeof msg, "device slot %d (CUDA device %d) failed and was removed from the round-robin", di, d.dev);
        publish_error(ctx, msg);
    }
    return lost;
}
std::vector<int> healthy_devs(j2k_ctx* ctx) {
    std::vector<int> h;
    for (int di = 0; di < (int)ctx->devs.size(); di++)
        if (!ctx->devs[di].failed) h.push_back(di);
    return h;
}

// Frames [f0, f1) over the healthy devices in contiguous blocks.  Blocking calls re-run the block of a device that was lost on
// the devices that are left; asynchronous submissions report the error (the device is out for the next call).
int run_frames(j2k_ctx* ctx, const HostJob& J, int f0, int f1, bool wait, std::vector<int>* used, bool timing, int* timing_dev) {
    const std::vector<int> H = healthy_devs(ctx);
    if (H.empty()) return fail(J2K_ERR_CUDA, "no usable device left in this context (every device failed)");
    const int n = f1 - f0, nh = (int)H.size();
    const int per = (n + nh - 1) / nh;
    struct Part { int di, a, b, rc; };
    std::vector<Part> parts;
    for (int k = 0; k < nh; k++) {
        const int a = f0 + k * per, b = a + per > f1 ? f1 : a + per;
        if (a >= b) break;
        Part pt{H[k], a, b, 0};
        if (fault_slot() == pt.di) pt.rc = fail(J2K_ERR_CUDA, "injected fault on device slot %d (J2K_FAULT_DEVICE)", pt.di);
        else pt.rc = enqueue_host_job(ctx, pt.di, J, a, b, timing && k == 0);
        if (timing && k == 0 && timing_dev) *timing_dev = pt.di;
        if (used) used->push_back(pt.di);
        parts.push_back(pt);
        if (pt.rc && !wait) break;
    }
    if (!wait) {
        for (auto& pt : parts)
            if (pt.rc) {
                // no ticket will exist for this submission: nothing of it may stay in flight on the caller's buffers
                for (auto& q : parts)
                    if (q.rc == 0 && !ctx->devs[q.di].failed && fault_slot() != q.di) (void)sync_dev(ctx, q.di);
                device_lost(ctx, pt.di);
                return pt.rc;
            }
        return 0;
    }
    for (auto& pt : parts) {
        if (ctx->devs[pt.di].failed || fault_slot() == pt.di) continue;
        const int r2 = sync_dev(ctx, pt.di);
        if (pt.rc == 0) pt.rc = r2;
    }
    int rc = 0;
    for (auto& pt : parts) {
        if (!pt.rc) continue;
        if (!device_lost(ctx, pt.di)) { if (!rc) rc = pt.rc; continue; }  // not the device: the caller's error
        if (timing_dev && *timing_dev == pt.di) *timing_dev = -1;
        const int r2 = run_frames(ctx, J, pt.a, pt.b, true, nullptr, false, nullptr);   // the lost device's block, on the others
        if (!rc) rc = r2;
    }
    return rc;
}

// Shards nframes over the devices in contiguous blocks (no collective), optionally waits.
int run_host_batch(j2k_ctx* ctx, const HostJob& J0, int nframes, bool wait, std::vector<int>* used) {
    HostJob J = J0;
    J.may_stage = wait;
    std::lock_guard<std::mutex> lk(ctx->mu);
    long long l0 = ctx->launches.load();
    int tdev = -1;
    int rc = run_frames(ctx, J, 0, nframes, wait, used, wait, &tdev);
    if (!wait) return rc;
    if (rc == 0 && tdev >= 0) {
        DeviceCtx& d = ctx->devs[tdev];
        if (set_dev(ctx, tdev)) return 0;
        j2k_timing t{};
        cudaEventElapsedTime(&t.h2d_ms, d.ev_t[0], d.ev_t[1]);
        cudaEventElapsedTime(&t.kernel_ms, d.ev_t[1], d.ev_t[2]);
        cudaEventElapsedTime(&t.d2h_ms, d.ev_t[2], d.ev_t[3]);
        cudaEventElapsedTime(&t.total_ms, d.ev_t[0], d.ev_t[3]);
        t.kernel_launches = (int32_t)(ctx->launches.load() - l0);
        ctx->last = t;
    }
    return rc;
}

size_t pix_bytes(int w, int h, int C, int B) { return (size_t)w * h * C * ((B + 7) / 8); }

// small synchronous helper for the package-API entry points: upload n buffers, run `body`, download
struct ApiIo { const void* h_in; void* h_out; size_t bytes; };

}  // namespace

// =================================================================== C ABI

#pragma GCC visibility push(default)
extern "C" {

int j2k_abi_version(void) { return J2K_B200_ABI_VERSION; }

// With a context: the message of the most recent failing call ON THAT CONTEXT, whichever thread made it (copied into a
// per-thread buffer so that the pointer stays valid); without: the calling thread's last message (j2k_init, size helpers).
const char* j2k_last_error(j2k_ctx* ctx) {
    if (!ctx) return t_err.c_str();
    thread_local std::string ret;
    std::lock_guard<std::mutex> lk(ctx->err_mu);
    ret = ctx->last_err;
    return ret.c_str();
}

// Copy form for bindings that must not hold a C pointer (cgo): writes at most cap - 1 bytes + NUL, returns the full length.
size_t j2k_last_error_copy(j2k_ctx* ctx, char* buf, size_t cap) {
    std::string msg;
    if (ctx) { std::lock_guard<std::mutex> lk(ctx->err_mu); msg = ctx->last_err; } else msg = t_err;
    if (buf && cap) {
        size_t n = msg.size() < cap - 1 ? msg.size() : cap - 1;
        memcpy(buf, msg.data(), n);
        buf[n] = 0;
    }
    return msg.size();
}

int j2k_init(j2k_ctx** out, const int* devices, int n_devices) {
    if (!out) return fail(J2K_ERR_INVALID_ARG, "ctx out pointer is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return fail(J2K_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    std::vector<int> ids;
    if (n_devices <= 0) {
        const char* env = getenv("J2K_B200_DEVICE");
        ids.push_back(env ? atoi(env) : 0);
    } else {
        for (int i = 0; i < n_devices; i++) ids.push_back(devices ? devices[i] : i);
    }
    for (int id : ids) if (id < 0 || id >= count) return fail(J2K_ERR_INVALID_ARG, "device %d out of range (0..%d)", id, count - 1);
    std::unique_ptr<j2k_ctx> ctx(new j2k_ctx());
    ctx->devs.resize(ids.size());
    for (size_t i = 0; i < ids.size(); i++) {
        DeviceCtx& d = ctx->devs[i];
        d.dev = ids[i];
        CK(cudaSetDevice(d.dev));
        CK(cudaStreamCreateWithFlags(&d.s_main, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&d.s_h2d, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&d.s_d2h, cudaStreamNonBlocking));
        for (auto& ev : d.ev_t) CK(cudaEventCreate(&ev));
        for (int k = 0; k < 2; k++) {
            CK(cudaEventCreateWithFlags(&d.ev_in[k], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&d.ev_k[k], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&d.ev_out[k], cudaEventDisableTiming));
        }
    }
    *out = ctx.release();
    return J2K_OK;
}

void j2k_shutdown(j2k_ctx* ctx) {
    if (!ctx) return;
    for (auto& d : ctx->devs) {
        cudaSetDevice(d.dev);
        cudaStreamSynchronize(d.s_main); cudaStreamSynchronize(d.s_h2d); cudaStreamSynchronize(d.s_d2h);
        d.plans.clear();
        for (auto& b : d.in) b.release();
        for (auto& b : d.out) b.release();
        for (auto& b : d.planes) b.release();
        for (auto& b : d.blk) b.release();
        for (auto& b : d.nbp) b.release();
        for (auto& b : d.gsh) b.release();
        for (auto& b : d.gmk) b.release();
        d.block_tables.clear();
        for (auto& b : d.api) b.release();
        for (auto& ev : d.ev_t) if (ev) cudaEventDestroy(ev);
        for (int k = 0; k < 2; k++) { cudaEventDestroy(d.ev_in[k]); cudaEventDestroy(d.ev_k[k]); cudaEventDestroy(d.ev_out[k]); }
        if (d.ev_ht) cudaEventDestroy(d.ev_ht);
        for (int k = 0; k < 2; k++) if (d.he_pin[k]) cudaFreeHost(d.he_pin[k]);
        cudaStreamDestroy(d.s_main); cudaStreamDestroy(d.s_h2d); cudaStreamDestroy(d.s_d2h);
    }
    pool_stop(ctx);
    for (auto& d : ctx->devs) { cudaSetDevice(d.dev); stager_release(d); }
    for (auto& kv : ctx->tickets) for (auto& t : kv.second) cudaEventDestroy(t.ev);
    for (void* p : ctx->pinned) cudaFreeHost(p);
    for (cudaEvent_t e : ctx->prof_ev) cudaEventDestroy(e);
    delete ctx;
}

int j2k_device_count(const j2k_ctx* ctx) { return ctx ? (int)ctx->devs.size() : 0; }
int j2k_device_failed(const j2k_ctx* ctx, int slot) {
    if (!ctx || slot < 0 || slot >= (int)ctx->devs.size()) return -1;
    return ctx->devs[slot].failed ? 1 : 0;
}
// CUDA devices visible to this process (0 when there is none or the runtime fails): what a binding passes to j2k_init to
// build a context over every GPU.
int j2k_visible_devices(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n < 0 ? 0 : n;
}
int64_t j2k_launch_count(const j2k_ctx* ctx) { return ctx ? (int64_t)ctx->launches.load() : 0; }

int j2k_last_timing(j2k_ctx* ctx, j2k_timing* out) {
    CtxGuard cg_(ctx);
    if (!ctx || !out) return fail(J2K_ERR_INVALID_ARG, "NULL argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    *out = ctx->last;
    return J2K_OK;
}

int j2k_set_profiling(j2k_ctx* ctx, int enabled) {
    CtxGuard cg_(ctx);
    if (!ctx) return fail(J2K_ERR_INVALID_ARG, "context is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->profiling = enabled != 0;
    return J2K_OK;
}

int j2k_get_profile(j2k_ctx* ctx, float* ms, int32_t* levels, int max) {
    CtxGuard cg_(ctx);
    if (!ctx || !ms || max < 0) return fail(J2K_ERR_INVALID_ARG, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    int n = (int)ctx->prof_level.size();
    if (n == 0) return 0;
    CK(cudaStreamSynchronize(ctx->prof_stream));
    for (int i = 0; i < n && i < max; i++) {
        CK(cudaEventElapsedTime(&ms[i], ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]));
        if (levels) levels[i] = ctx->prof_level[i];
    }
    return n;
}

void* j2k_acquire_buffer(j2k_ctx* ctx, size_t nbytes) {
    CtxGuard cg_(ctx);
    if (!ctx || nbytes == 0) { fail(J2K_ERR_INVALID_ARG, "NULL context or zero size"); return nullptr; }
    void* p = nullptr;
    cudaSetDevice(ctx->devs[0].dev);
    cudaError_t e = cudaHostAlloc(&p, nbytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { fail(J2K_ERR_NOMEM, "cudaHostAlloc(%zu) failed: %s", nbytes, cudaGetErrorString(e)); return nullptr; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->pinned.push_back(p);
    return p;
}

void j2k_release_buffer(j2k_ctx* ctx, void* p) {
    if (!ctx || !p) return;
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (size_t i = 0; i < ctx->pinned.size(); i++)
        if (ctx->pinned[i] == p) { ctx->pinned.erase(ctx->pinned.begin() + (long)i); cudaFreeHost(p); return; }
}

// ---- sizes

size_t j2k_fwd_pixel_bytes(const j2k_fwd_params* p) { return p ? pix_bytes(p->width, p->height, p->components, p->bit_depth) : 0; }
size_t j2k_fwd_coeff_count(const j2k_fwd_params* p) { return p ? (size_t)p->width * p->height * p->components : 0; }
size_t j2k_inv_pixel_bytes(const j2k_inv_params* p) { return p ? pix_bytes(p->xsiz - p->xosiz, p->ysiz - p->yosiz, p->components, p->bit_depth) : 0; }
size_t j2k_inv_coeff_count(const j2k_inv_params* p) { return p ? (size_t)(p->xsiz - p->xosiz) * (p->ysiz - p->yosiz) * p->components : 0; }

int j2k_fwd_tile_bounds(const j2k_fwd_params* p, int idx, int32_t b[4]) {
    if (!p) return fail(J2K_ERR_INVALID_ARG, "params is NULL");
    std::vector<TileGeom> t;
    int n = fwd_tiles(p, &t);
    if (b && idx >= 0 && idx < n) { b[0] = t[idx].x0; b[1] = t[idx].y0; b[2] = t[idx].x0 + t[idx].tw; b[3] = t[idx].y0 + t[idx].th; }
    return n;
}
int j2k_inv_tile_bounds(const j2k_inv_params* p, int idx, int32_t b[4]) {
    if (!p) return fail(J2K_ERR_INVALID_ARG, "params is NULL");
    if (p->xtsiz <= 0 || p->ytsiz <= 0) return fail(J2K_ERR_INVALID_ARG, "invalid tile size");
    std::vector<TileGeom> t;
    int n = inv_tiles(p, &t);
    if (b && idx >= 0 && idx < n) { b[0] = t[idx].x0; b[1] = t[idx].y0; b[2] = t[idx].x0 + t[idx].tw; b[3] = t[idx].y0 + t[idx].th; }
    return n;
}

// ---- forward

static int forward_host(j2k_ctx* ctx, const j2k_fwd_params* p, int nframes, const void* pixels, size_t stride, int32_t* coeffs_out,
                        bool planar, bool wait, std::vector<int>* used) {
    CtxGuard cg_(ctx);
    if (!ctx) return fail(J2K_ERR_INVALID_ARG, "context is NULL");
    Spec s;
    int rc = spec_from_fwd(p, planar, s);
    if (rc) return rc;
    if (!pixels || !coeffs_out) return fail(J2K_ERR_INVALID_ARG, "NULL buffer");
    if (nframes <= 0) return fail(J2K_ERR_INVALID_ARG, "nframes must be positive");
    HostJob J{};
    J.fwd = true; J.planar = planar; J.spec = &s; J.pblob = p; J.pbytes = sizeof *p;
    J.h_pix_in = (const unsigned char*)pixels; J.h_coef_out = coeffs_out;
    J.pix_bytes_per_frame = j2k_fwd_pixel_bytes(p); J.coeffs_per_frame = (long long)j2k_fwd_coeff_count(p);
    J.frame_stride_bytes = planar ? 0 : stride;
    if (!planar && stride < J.pix_bytes_per_frame) return fail(J2K_ERR_SIZE, "frame stride %zu smaller than a frame (%zu bytes)", stride, J.pix_bytes_per_frame);
    return run_host_batch(ctx, J, nframes, wait, used);
}

int j2k_forward(j2k_ctx* ctx, const j2k_fwd_params* p, const void* pixels, size_t nbytes, int32_t* coeffs_out, size_t ncoeffs) {
    CtxGuard cg_(ctx);
    if (!p) return fail(J2K_ERR_INVALID_ARG, "params is NULL");
    size_t need = j2k_fwd_pixel_bytes(p);
    if (nbytes < need) return fail(J2K_ERR_SIZE, "insufficient pixel data: got %zu bytes, need %zu", nbytes, need);  // encoder.go:346-348
    if (ncoeffs < j2k_fwd_coeff_count(p)) return fail(J2K_ERR_SIZE, "coefficient buffer too small: got %zu, need %zu", ncoeffs, j2k_fwd_coeff_count(p));
    return forward_host(ctx, p, 1, pixels, need, coeffs_out, false, true, nullptr);
}

int j2k_forward_planar(j2k_ctx* ctx, const j2k_fwd_params* p, const int32_t* const* planes, int32_t* coeffs_out, size_t ncoeffs) {
    CtxGuard cg_(ctx);
    if (!p || !planes) return fail(J2K_ERR_INVALID_ARG, "NULL argument");
    int rc = validate_common(p->width, p->height, p->components, p->bit_depth, p->num_levels);
    if (rc) return rc;
    if (ncoeffs < j2k_fwd_coeff_count(p)) return fail(J2K_ERR_SIZE, "coefficient buffer too small");
    size_t hw = (size_t)p->width * p->height;
    std::vector<int32_t> packed(hw * p->components);  // gather the component slices (they are separate Go slices)
    for (int c = 0; c < p->components; c++) {
        if (!planes[c]) return fail(J2K_ERR_INVALID_ARG, "component %d is NULL", c);
        memcpy(packed.data() + c * hw, planes[c], hw * 4);
    }
    return forward_host(ctx, p, 1, packed.data(), 0, coeffs_out, true, true, nullptr);
}

// The same with the component planes in ONE buffer (component c at planes + c * plane_stride): the form a cgo caller can
// use, because a [][]int32 (Go pointers to Go pointers) cannot cross the boundary.
int j2k_forward_planar_flat(j2k_ctx* ctx, const j2k_fwd_params* p, const int32_t* planes, size_t plane_stride, int32_t* coeffs_out,
                            size_t ncoeffs) {
    CtxGuard cg_(ctx);
    if (!p || !planes) return fail(J2K_ERR_INVALID_ARG, "NULL argument");
    int rc = validate_common(p->width, p->height, p->components, p->bit_depth, p->num_levels);
    if (rc) return rc;
    const size_t hw = (size_t)p->width * p->height;
    if (plane_stride < hw) return fail(J2K_ERR_SIZE, "plane stride %zu smaller than a plane (%zu samples)", plane_stride, hw);
    if (ncoeffs < j2k_fwd_coeff_count(p)) return fail(J2K_ERR_SIZE, "coefficient buffer too small");
    if (plane_stride == hw) return forward_host(ctx, p, 1, planes, 0, coeffs_out, true, true, nullptr);
    std::vector<int32_t> packed(hw * p->components);
    for (int c = 0; c < p->components; c++) memcpy(packed.data() + c * hw, planes + c * plane_stride, hw * 4);
    return forward_host(ctx, p, 1, packed.data(), 0, coeffs_out, true, true, nullptr);
}

int j2k_forward_batch(j2k_ctx* ctx, const j2k_fwd_params* p, int nframes, const void* pixels, size_t frame_stride_bytes, int32_t* coeffs_out) {
    CtxGuard cg_(ctx);
    return forward_host(ctx, p, nframes, pixels, frame_stride_bytes, coeffs_out, false, true, nullptr);
}

int j2k_forward_device(j2k_ctx* ctx, int dev, const j2k_fwd_params* p, int nframes, const void* d_pixels, size_t frame_stride_bytes,
                       int32_t* d_coeffs, void* cuda_stream) {
    CtxGuard cg_(ctx);
    int rc = set_dev(ctx, dev);
    if (rc) return rc;
    Spec s;
    if ((rc = spec_from_fwd(p, false, s))) return rc;
    if (!d_pixels || !d_coeffs || nframes <= 0) return fail(J2K_ERR_INVALID_ARG, "bad device buffers / nframes");
    int bps = s.bit_depth <= 8 ? 1 : 2;
    if (frame_stride_bytes < j2k_fwd_pixel_bytes(p) || frame_stride_bytes % bps) return fail(J2K_ERR_SIZE, "bad frame stride");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceCtx& d = ctx->devs[dev];
    Plan* P = nullptr;
    if ((rc = get_plan(d, s, p, sizeof *p, nframes, (long long)(frame_stride_bytes / bps), &P, cuda_stream))) return rc;
    rc = run_plan(ctx, *P, (void*)d_pixels, d_coeffs, nullptr, false, cuda_stream ? (cudaStream_t)cuda_stream : d.s_main);
    return rc < 0 ? rc : J2K_OK;
}

// ---- inverse

static int inverse_host(j2k_ctx* ctx, const j2k_inv_params* p, int nframes, const int32_t* coeffs_in, void* pixels_out, size_t stride,
                        int32_t* planes_out, bool wait, std::vector<int>* used) {
    CtxGuard cg_(ctx);
    if (!ctx) return fail(J2K_ERR_INVALID_ARG, "context is NULL");
    Spec s;
    int rc = spec_from_inv(p, planes_out != nullptr, s);
    if (rc) return rc;
    if (!coeffs_in || !pixels_out) return fail(J2K_ERR_INVALID_ARG, "NULL buffer");
    if (nframes <= 0) return fail(J2K_ERR_INVALID_ARG, "nframes must be positive");
    HostJob J{};
    J.fwd = false; J.planar = false; J.spec = &s; J.pblob = p; J.pbytes = sizeof *p;
    J.h_coef_in = coeffs_in; J.h_pix_out = (unsigned char*)pixels_out; J.h_planes = planes_out;
    J.pix_bytes_per_frame = j2k_inv_pixel_bytes(p); J.coeffs_per_frame = (long long)j2k_inv_coeff_count(p);
    J.frame_stride_bytes = stride;
    if (stride < J.pix_bytes_per_frame) return fail(J2K_ERR_SIZE, "frame stride smaller than a frame");
    return run_host_batch(ctx, J, nframes, wait, used);
}

int j2k_inverse(j2k_ctx* ctx, const j2k_inv_params* p, const int32_t* coeffs_in, size_t ncoeffs, void* pixels_out, size_t nbytes,
                int32_t* planes_out) {
    CtxGuard cg_(ctx);
    if (!p) return fail(J2K_ERR_INVALID_ARG, "params is NULL");
    if (ncoeffs < j2k_inv_coeff_count(p)) return fail(J2K_ERR_SIZE, "coefficient buffer too small: got %zu, need %zu", ncoeffs, j2k_inv_coeff_count(p));
    size_t need = j2k_inv_pixel_bytes(p);
    if (nbytes < need) return fail(J2K_ERR_SIZE, "pixel buffer too small: got %zu bytes, need %zu", nbytes, need);
    return inverse_host(ctx, p, 1, coeffs_in, pixels_out, need, planes_out, true, nullptr);
}

int j2k_inverse_batch(j2k_ctx* ctx, const j2k_inv_params* p, int nframes, const int32_t* coeffs_in, void* pixels_out,
                      size_t frame_stride_bytes, int32_t* planes_out) {
    CtxGuard cg_(ctx);
    return inverse_host(ctx, p, nframes, coeffs_in, pixels_out, frame_stride_bytes, planes_out, true, nullptr);
}

int j2k_inverse_device(j2k_ctx* ctx, int dev, const j2k_inv_params* p, int nframes, const int32_t* d_coeffs, void* d_pixels,
                       size_t frame_stride_bytes, int32_t* d_planes, void* cuda_stream) {
    CtxGuard cg_(ctx);
    int rc = set_dev(ctx, dev);
    if (rc) return rc;
    Spec s;
    if ((rc = spec_from_inv(p, d_planes != nullptr, s))) return rc;
    if (!d_pixels || !d_coeffs || nframes <= 0) return fail(J2K_ERR_INVALID_ARG, "bad device buffers / nframes");
    int bps = s.bit_depth <= 8 ? 1 : 2;
    if (frame_stride_bytes < j2k_inv_pixel_bytes(p) || frame_stride_bytes % bps) return fail(J2K_ERR_SIZE, "bad frame stride");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceCtx& d = ctx->devs[dev];
    Plan* P = nullptr;
    if ((rc = get_plan(d, s, p, sizeof *p, nframes, (long long)(frame_stride_bytes / bps), &P, cuda_stream))) return rc;
    rc = run_plan(ctx, *P, d_pixels, (void*)d_coeffs, d_planes, false, cuda_stream ? (cudaStream_t)cuda_stream : d.s_main);
    return rc < 0 ? rc : J2K_OK;
}

// ---- code-block interface

int j2k_codeblock_layout(int width, int height, int num_levels, int cb_width, int cb_height, j2k_cblk* out, int max_blocks) {
    if (width <= 0 || height <= 0 || num_levels < 0 || num_levels > J2K_MAX_LEVELS) return fail(J2K_ERR_INVALID_ARG, "invalid plane geometry");
    int rc = validate_cb(cb_width, cb_height);
    if (rc) return rc;
    std::vector<j2k_cblk> t;
    codeblock_layout(width, height, num_levels, cb_width, cb_height, t);
    if (out)
        for (size_t i = 0; i < t.size() && (int)i < max_blocks; i++) out[i] = t[i];
    return (int)t.size();
}

size_t j2k_fwd_block_count(const j2k_fwd_params* p, int cb_width, int cb_height) {
    if (!p || validate_cb(cb_width, cb_height)) return 0;
    std::vector<TileGeom> tiles;
    fwd_tiles(p, &tiles);
    return frame_block_count(tiles, p->components, p->num_levels, cb_width, cb_height);
}

size_t j2k_inv_block_count(const j2k_inv_params* p, int cb_width, int cb_height) {
    if (!p || validate_cb(cb_width, cb_height)) return 0;
    std::vector<TileGeom> tiles;
    inv_tiles(p, &tiles);
    return frame_block_count(tiles, p->components, p->num_levels, cb_width, cb_height);
}

// In block mode the coefficients leave exactly as encodeCodeBlock hands them to T1 (encoder.go:3294-3300): the << 6 of
// the classic lossless path is applied once, inside the forward kernel (the quantizer stage does it for free).
static j2k_fwd_params block_mode_params(const j2k_fwd_params* p) {
    j2k_fwd_params q = *p;
    q.fuse_t1_shift = (q.reversible && !q.htj2k) ? 1 : 0;
    return q;
}

int j2k_forward_blocks(j2k_ctx* ctx, const j2k_fwd_params* p, int cb_width, int cb_height, int nframes, const void* pixels,
                       size_t frame_stride_bytes, int32_t* blocks_out, int32_t* numbps_out) {
    CtxGuard cg_(ctx);
    if (!ctx) return fail(J2K_ERR_INVALID_ARG, "context is NULL");
    if (!p) return fail(J2K_ERR_INVALID_ARG, "params is NULL");
    int rc = validate_cb(cb_width, cb_height);
    if (rc) return rc;
    const j2k_fwd_params q = block_mode_params(p);
    Spec s;
    if ((rc = spec_from_fwd(&q, false, s))) return rc;
    if (!pixels || !blocks_out) return fail(J2K_ERR_INVALID_ARG, "NULL buffer");
    if (nframes <= 0) return fail(J2K_ERR_INVALID_ARG, "nframes must be positive");
    HostJob J{};
    J.fwd = true; J.planar = false; J.spec = &s; J.pblob = &q; J.pbytes = sizeof q;
    J.h_pix_in = (const unsigned char*)pixels; J.h_coef_out = blocks_out; J.h_numbps = numbps_out;
    J.pix_bytes_per_frame = j2k_fwd_pixel_bytes(&q); J.coeffs_per_frame = (long long)j2k_fwd_coeff_count(&q);
    J.frame_stride_bytes = frame_stride_bytes; J.cb_w = cb_width; J.cb_h = cb_height;
    if (frame_stride_bytes < J.pix_bytes_per_frame) return fail(J2K_ERR_SIZE, "frame stride %zu smaller than a frame (%zu bytes)", frame_stride_bytes, J.pix_bytes_per_frame);
    return run_host_batch(ctx, J, nframes, true, nullptr);
}

int j2k_inverse_blocks(j2k_ctx* ctx, const j2k_inv_params* p, int cb_width, int cb_height, int nframes, const int32_t* blocks_in,
                       void* pixels_out, size_t frame_stride_bytes, int32_t* planes_out) {
    CtxGuard cg_(ctx);
    return j2k_inverse_blocks_roi(ctx, p, cb_width, cb_height, nframes, blocks_in, nullptr, pixels_out, frame_stride_bytes, planes_out);
}

int j2k_inverse_blocks_roi(j2k_ctx* ctx, const j2k_inv_params* p, int cb_width, int cb_height, int nframes, const int32_t* blocks_in,
                           const int32_t* roi_maxshift, void* pixels_out, size_t frame_stride_bytes, int32_t* planes_out) {
    CtxGuard cg_(ctx);
    return j2k_inverse_blocks_roi_general(ctx, p, cb_width, cb_height, nframes, blocks_in, roi_maxshift, nullptr, nullptr, pixels_out,
                                          frame_stride_bytes, planes_out);
}

int j2k_inverse_blocks_roi_general(j2k_ctx* ctx, const j2k_inv_params* p, int cb_width, int cb_height, int nframes, const int32_t* blocks_in,
                                   const int32_t* roi_maxshift, const int32_t* block_scale_shift, const uint8_t* sample_mask, void* pixels_out,
                                   size_t frame_stride_bytes, int32_t* planes_out) {
    CtxGuard cg_(ctx);
    if (!ctx) return fail(J2K_ERR_INVALID_ARG, "context is NULL");
    if (!p) return fail(J2K_ERR_INVALID_ARG, "params is NULL");
    int rc = validate_cb(cb_width, cb_height);
    if (rc) return rc;
    Spec s;
    if ((rc = spec_from_inv(p, planes_out != nullptr, s))) return rc;
    if (!blocks_in || !pixels_out) return fail(J2K_ERR_INVALID_ARG, "NULL buffer");
    if (nframes <= 0) return fail(J2K_ERR_INVALID_ARG, "nframes must be positive");
    HostJob J{};
    J.fwd = false; J.planar = false; J.spec = &s; J.pblob = p; J.pbytes = sizeof *p;
    J.h_coef_in = blocks_in; J.h_pix_out = (unsigned char*)pixels_out; J.h_planes = planes_out;
    J.pix_bytes_per_frame = j2k_inv_pixel_bytes(p); J.coeffs_per_frame = (long long)j2k_inv_coeff_count(p);
    J.frame_stride_bytes = frame_stride_bytes; J.cb_w = cb_width; J.cb_h = cb_height;
    if ((rc = make_roi(roi_maxshift, p->components, J.roi))) return rc;
    if (frame_stride_bytes < J.pix_bytes_per_frame) return fail(J2K_ERR_SIZE, "frame stride smaller than a frame");
    if (sample_mask && !block_scale_shift) return fail(J2K_ERR_INVALID_ARG, "a sample mask needs the per-block general-scaling shifts");
    if (block_scale_shift) {
        const size_t nblk = j2k_inv_block_count(p, cb_width, cb_height);
        for (size_t i = 0; i < nblk * (size_t)nframes; i++)
            if (block_scale_shift[i] < 0 || block_scale_shift[i] > 30)
                return fail(J2K_ERR_INVALID_ARG, "invalid general-scaling shift %d for block %zu (0..30)", block_scale_shift[i], i);
        J.h_block_shift = block_scale_shift;
        J.h_sample_mask = sample_mask;
    }
    return run_host_batch(ctx, J, nframes, true, nullptr);
}

int j2k_gather_blocks_device(j2k_ctx* ctx, int dev, const j2k_fwd_params* p, int cb_width, int cb_height, int nframes,
                             const int32_t* d_coeffs, int32_t* d_blocks, int32_t* d_numbps, void* cuda_stream) {
    CtxGuard cg_(ctx);
    int rc = set_dev(ctx, dev);
    if (rc) return rc;
    if (!p) return fail(J2K_ERR_INVALID_ARG, "params is NULL");
    if ((rc = validate_cb(cb_width, cb_height))) return rc;
    Spec s;
    if ((rc = spec_from_fwd(p, false, s))) return rc;
    if (!d_coeffs || !d_blocks || nframes <= 0) return fail(J2K_ERR_INVALID_ARG, "bad device buffers / nframes");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceCtx& d = ctx->devs[dev];
    BlockTable* BT = nullptr;
    if ((rc = get_block_table(d, s, cb_width, cb_height, (long long)j2k_fwd_coeff_count(p), &BT))) return rc;
    // cblkNumbps counts bit planes above the 6 fractional bits of the T1 fixed point (encoder.go:3349-3362).  The coefficients
    // carry those bits when they come from the 9/7 quantizer (scale 64) or from a 5/3 forward with fuse_t1_shift; a classic 5/3
    // plane WITHOUT the fused shift holds plain integers (the caller shifts the block copies, encoder.go:3294-3300), whose bit
    // length already is the count.  HTJ2K has no fixed point.
    const bool sub6 = !s.htj2k && !(s.reversible && !s.fuse_shift);
    return launch_gather(ctx, *BT, nframes, d_coeffs, d_blocks, d_numbps, sub6, cuda_stream ? (cudaStream_t)cuda_stream : d.s_main);
}

int j2k_scatter_blocks_device(j2k_ctx* ctx, int dev, const j2k_inv_params* p, int cb_width, int cb_height, int nframes,
                              const int32_t* d_blocks, int32_t* d_coeffs, void* cuda_stream) {
    CtxGuard cg_(ctx);
    return j2k_scatter_blocks_roi_device(ctx, dev, p, cb_width, cb_height, nframes, d_blocks, nullptr, d_coeffs, cuda_stream);
}

int j2k_scatter_blocks_roi_device(j2k_ctx* ctx, int dev, const j2k_inv_params* p, int cb_width, int cb_height, int nframes,
                                  const int32_t* d_blocks, const int32_t* roi_maxshift, int32_t* d_coeffs, void* cuda_stream) {
    CtxGuard cg_(ctx);
    return j2k_scatter_blocks_roi_general_device(ctx, dev, p, cb_width, cb_height, nframes, d_blocks, roi_maxshift, nullptr, nullptr, d_coeffs,
                                                 cuda_stream);
}

int j2k_scatter_blocks_roi_general_device(j2k_ctx* ctx, int dev, const j2k_inv_params* p, int cb_width, int cb_height, int nframes,
                                          const int32_t* d_blocks, const int32_t* roi_maxshift, const int32_t* d_block_scale_shift,
                                          const uint8_t* d_sample_mask, int32_t* d_coeffs, void* cuda_stream) {
    CtxGuard cg_(ctx);
    int rc = set_dev(ctx, dev);
    if (rc) return rc;
    if (!p) return fail(J2K_ERR_INVALID_ARG, "params is NULL");
    if ((rc = validate_cb(cb_width, cb_height))) return rc;
    Spec s;
    if ((rc = spec_from_inv(p, false, s))) return rc;
    if (!d_coeffs || !d_blocks || nframes <= 0) return fail(J2K_ERR_INVALID_ARG, "bad device buffers / nframes");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceCtx& d = ctx->devs[dev];
    BlockTable* BT = nullptr;
    if ((rc = get_block_table(d, s, cb_width, cb_height, (long long)j2k_inv_coeff_count(p), &BT))) return rc;
    RoiShifts roi;
    if ((rc = make_roi(roi_maxshift, p->components, roi))) return rc;
    if (d_sample_mask && !d_block_scale_shift) return fail(J2K_ERR_INVALID_ARG, "a sample mask needs the per-block general-scaling shifts");
    roi.block_shift = (const int*)d_block_scale_shift;   // values must lie in 0..30 (not checked: device memory)
    roi.sample_mask = d_sample_mask;
    return launch_scatter(ctx, *BT, nframes, d_blocks, d_coeffs, cuda_stream ? (cudaStream_t)cuda_stream : d.s_main, roi);
}

// ---- HTJ2K block decoding on the device (decode side)

static_assert(sizeof(j2k_ht_cblk) == 16 && sizeof(HtBlock) == 16, "j2k_ht_cblk and the kernel's record must agree");

namespace ht_host {
#define J2K_HT_TABLE static const
#include "j2k_ht_tables.inc"
#undef J2K_HT_TABLE
}  // namespace ht_host

int j2k_ht_table(int which, uint16_t* out) {
    const unsigned short* t = which == 0 ? ht_host::HT_VLC_TBL0 : which == 1 ? ht_host::HT_VLC_TBL1 : which == 2 ? ht_host::HT_UVLC_TBL0
                              : which == 3 ? ht_host::HT_UVLC_TBL1 : nullptr;
    if (!t) return fail(J2K_ERR_INVALID_ARG, "no such table: %d", which);
    const int n = which < 2 ? 1024 : which == 2 ? 320 : 256;
    if (out) memcpy(out, t, (size_t)n * 2);
    return n;
}

static int64_t make_ticket(j2k_ctx* ctx, const std::vector<int>& used);

// Argument checks shared by the HT entry points; fills the geometry they need.
static int ht_check(j2k_ctx* ctx, const j2k_inv_params* p, int cbw, int cbh, int nframes, const uint8_t* bytes, size_t nbytes,
                    const j2k_ht_cblk* cblks, bool want_planes, Spec& s, size_t& nblk) {
    if (!ctx) return fail(J2K_ERR_INVALID_ARG, "context is NULL");
    if (!p) return fail(J2K_ERR_INVALID_ARG, "params is NULL");
    int rc = validate_ht_cb(cbw, cbh);
    if (rc) return rc;
    if ((rc = spec_from_inv(p, want_planes, s))) return rc;
    if (!cblks || (!bytes && nbytes)) return fail(J2K_ERR_INVALID_ARG, "NULL buffer");
    if (nframes <= 0) return fail(J2K_ERR_INVALID_ARG, "nframes must be positive");
    nblk = j2k_inv_block_count(p, cbw, cbh);
    for (size_t i = 0; i < nblk * (size_t)nframes; i++)
        if (cblks[i].length && (cblks[i].offset > nbytes || cblks[i].length > nbytes - cblks[i].offset))
            return fail(J2K_ERR_SIZE, "code-block %zu: segment [%llu, +%u) lies outside the %zu-byte stream", i,
                        (unsigned long long)cblks[i].offset, cblks[i].length, nbytes);
    return 0;
}

// HTDecoder.Decode for every block, block-major planes back to the host (the code-block interface's layout): cleanup segments
// and records up, one launch, planes down.  Frames are sharded over the context's devices in contiguous blocks.
int j2k_ht_decode_blocks(j2k_ctx* ctx, const j2k_inv_params* p, int cb_width, int cb_height, int nframes, const uint8_t* bytes,
                         size_t nbytes, const j2k_ht_cblk* cblks, int32_t* blocks_out, int32_t* status_out) {
    CtxGuard cg_(ctx);
    Spec s;
    size_t nblk = 0;
    int rc = ht_check(ctx, p, cb_width, cb_height, nframes, bytes, nbytes, cblks, false, s, nblk);
    if (rc) return rc;
    if (!blocks_out) return fail(J2K_ERR_INVALID_ARG, "NULL buffer");
    const long long cpf = (long long)j2k_inv_coeff_count(p);
    std::lock_guard<std::mutex> lk(ctx->mu);
    const std::vector<int> H = healthy_devs(ctx);   // devices lost earlier stay out (j2k_device_failed)
    if (H.empty()) return fail(J2K_ERR_CUDA, "no usable device left in this context (every device failed)");
    const int nd = (int)H.size();
    const int per = (nframes + nd - 1) / nd;
    std::vector<std::vector<HtBlock>> recs(ctx->devs.size());   // re-based records: alive until the devices are synchronised below
    int used = 0;
    for (int hk = 0; hk < nd && rc == 0; hk++) {
        const int di = H[hk];
        const int f0 = hk * per, f1 = f0 + per > nframes ? nframes : f0 + per;
        if (f0 >= f1) break;
        used = hk + 1;
        if ((rc = sync_dev(ctx, di))) break;   // the slot buffers may still serve an asynchronous job
        DeviceCtx& d = ctx->devs[di];
        const int n = f1 - f0;
        const size_t count = (size_t)n * nblk;
        const j2k_ht_cblk* src = cblks + (size_t)f0 * nblk;
        unsigned long long lo = ~0ull, hi = 0;
        for (size_t i = 0; i < count; i++)
            if (src[i].length) {
                if (src[i].offset < lo) lo = src[i].offset;
                if (src[i].offset + src[i].length > hi) hi = src[i].offset + src[i].length;
            }
        if (hi == 0) lo = 0;
        std::vector<HtBlock>& r = recs[di];
        r.resize(count);
        for (size_t i = 0; i < count; i++) {
            r[i].offset = src[i].length ? src[i].offset - lo : 0;
            r[i].length = src[i].length; r[i].kmax = src[i].kmax; r[i].mmsb = src[i].missing_msbs; r[i].reserved = 0;
        }
        BlockTable* BT = nullptr;
        if ((rc = get_block_table(d, s, cb_width, cb_height, cpf, &BT))) break;
        if ((rc = d.ht_bytes[0].ensure((size_t)(hi - lo) + 16))) break;
        if ((rc = d.ht_desc[0].ensure(count * sizeof(HtBlock) + 16))) break;
        if (status_out && (rc = d.ht_status[0].ensure(count * 4 + 16))) break;
        if ((rc = d.blk[0].ensure((size_t)n * cpf * 4))) break;
        cudaStream_t st = d.s_main;
        if (hi > lo) CK(cudaMemcpyAsync(d.ht_bytes[0].p, bytes + lo, (size_t)(hi - lo), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d.ht_desc[0].p, r.data(), count * sizeof(HtBlock), cudaMemcpyHostToDevice, st));
        if ((rc = launch_ht_decode(ctx, d, *BT, cb_width, cb_height, n, (const unsigned char*)d.ht_bytes[0].p, (const HtBlock*)d.ht_desc[0].p,
                                   (int32_t*)d.blk[0].p, 0, status_out ? (int32_t*)d.ht_status[0].p : nullptr, st))) break;
        if (status_out) CK(cudaMemcpyAsync(status_out + (size_t)f0 * nblk, d.ht_status[0].p, count * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(blocks_out + (size_t)f0 * cpf, d.blk[0].p, (size_t)n * cpf * 4, cudaMemcpyDeviceToHost, st));
    }
    for (int hk = 0; hk < used; hk++) {
        int r2 = sync_dev(ctx, H[hk]);
        if (rc == 0) rc = r2;
        if (r2) device_lost(ctx, H[hk]);
    }
    return rc;
}

// The decode tail on the device, on the pipelined host-batch path (sub-batches on three streams, pinned staging for pageable
// callers): cleanup segments + records up, HT decode into the coefficient planes, the inverse plan, pixels down.
static int inverse_ht_host(j2k_ctx* ctx, const j2k_inv_params* p, int cbw, int cbh, int nframes, const uint8_t* bytes, size_t nbytes,
                           const j2k_ht_cblk* cblks, void* pixels_out, size_t frame_stride_bytes, int32_t* planes_out, int32_t* status_out,
                           bool wait, std::vector<int>* used) {
    Spec s;
    size_t nblk = 0;
    int rc = ht_check(ctx, p, cbw, cbh, nframes, bytes, nbytes, cblks, planes_out != nullptr, s, nblk);
    if (rc) return rc;
    if (!pixels_out) return fail(J2K_ERR_INVALID_ARG, "NULL buffer");
    HostJob J{};
    J.fwd = false; J.planar = false; J.spec = &s; J.pblob = p; J.pbytes = sizeof *p;
    J.h_pix_out = (unsigned char*)pixels_out; J.h_planes = planes_out;
    J.pix_bytes_per_frame = j2k_inv_pixel_bytes(p); J.coeffs_per_frame = (long long)j2k_inv_coeff_count(p);
    J.frame_stride_bytes = frame_stride_bytes; J.cb_w = cbw; J.cb_h = cbh;
    J.h_ht_bytes = bytes; J.h_ht_recs = cblks; J.h_ht_status = status_out;
    if (frame_stride_bytes < J.pix_bytes_per_frame) return fail(J2K_ERR_SIZE, "frame stride smaller than a frame");
    return run_host_batch(ctx, J, nframes, wait, used);
}

int j2k_inverse_ht(j2k_ctx* ctx, const j2k_inv_params* p, int cb_width, int cb_height, int nframes, const uint8_t* bytes, size_t nbytes,
                   const j2k_ht_cblk* cblks, void* pixels_out, size_t frame_stride_bytes, int32_t* planes_out, int32_t* status_out) {
    CtxGuard cg_(ctx);
    return inverse_ht_host(ctx, p, cb_width, cb_height, nframes, bytes, nbytes, cblks, pixels_out, frame_stride_bytes, planes_out, status_out,
                           true, nullptr);
}

int64_t j2k_submit_inverse_ht(j2k_ctx* ctx, const j2k_inv_params* p, int cb_width, int cb_height, int nframes, const uint8_t* bytes,
                              size_t nbytes, const j2k_ht_cblk* cblks, void* pixels_out, size_t frame_stride_bytes, int32_t* planes_out,
                              int32_t* status_out) {
    CtxGuard cg_(ctx);
    std::vector<int> used;
    int rc = inverse_ht_host(ctx, p, cb_width, cb_height, nframes, bytes, nbytes, cblks, pixels_out, frame_stride_bytes, planes_out, status_out,
                             false, &used);
    if (rc) return rc;
    return make_ticket(ctx, used);
}

int j2k_ht_decode_device(j2k_ctx* ctx, int dev, const j2k_inv_params* p, int cb_width, int cb_height, int nframes, const uint8_t* d_bytes,
                         const j2k_ht_cblk* d_cblks, int32_t* d_out, int to_planes, int32_t* d_status, void* cuda_stream) {
    CtxGuard cg_(ctx);
    int rc = set_dev(ctx, dev);
    if (rc) return rc;
    if (!p) return fail(J2K_ERR_INVALID_ARG, "params is NULL");
    if ((rc = validate_ht_cb(cb_width, cb_height))) return rc;
    Spec s;
    if ((rc = spec_from_inv(p, false, s))) return rc;
    if (!d_bytes || !d_cblks || !d_out || nframes <= 0) return fail(J2K_ERR_INVALID_ARG, "bad device buffers / nframes");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceCtx& d = ctx->devs[dev];
    BlockTable* BT = nullptr;
    if ((rc = get_block_table(d, s, cb_width, cb_height, (long long)j2k_inv_coeff_count(p), &BT))) return rc;
    return launch_ht_decode(ctx, d, *BT, cb_width, cb_height, nframes, d_bytes, (const HtBlock*)d_cblks, d_out, to_planes ? 1 : 0, d_status,
                            cuda_stream ? (cudaStream_t)cuda_stream : d.s_main);
}

// ---- HTJ2K block encoding on the device (encode side)

namespace ht_host_enc {   // host copy of the device table (same generated file)
#define J2K_HT_TABLE static const
#include "j2k_ht_enc_tables.inc"
#undef J2K_HT_TABLE
}  // namespace ht_host_enc

int j2k_ht_enc_table(int which, uint16_t* out) {
    if (which != 0 && which != 1) return fail(J2K_ERR_INVALID_ARG, "no such table: %d", which);
    if (out) memcpy(out, which ? ht_host_enc::HT_ENC_TBL1 : ht_host_enc::HT_ENC_TBL0, 2048 * 2);
    return 2048;
}

size_t j2k_ht_encode_bound(const j2k_fwd_params* p, int cb_width, int cb_height, int kmax_max, int nframes) {
    if (!p || validate_ht_cb(cb_width, cb_height) || kmax_max <= 0 || kmax_max >= 31 || nframes <= 0) return 0;
    return ht_enc_bound(cb_width, cb_height, kmax_max, (long long)j2k_fwd_block_count(p, cb_width, cb_height) * nframes);
}

int j2k_ht_encode_device(j2k_ctx* ctx, int dev, const j2k_fwd_params* p, int cb_width, int cb_height, int nframes, const int32_t* d_coeffs,
                         const uint8_t* kmax, uint8_t* d_bytes, size_t bytes_cap, j2k_ht_cblk* d_cblks, uint64_t* d_offsets,
                         void* cuda_stream) {
    CtxGuard cg_(ctx);
    int rc = set_dev(ctx, dev);
    if (rc) return rc;
    if (!p) return fail(J2K_ERR_INVALID_ARG, "params is NULL");
    if ((rc = validate_ht_cb(cb_width, cb_height))) return rc;
    Spec s;
    if ((rc = spec_from_fwd(p, false, s))) return rc;
    if (!d_coeffs || !kmax || !d_bytes || !d_cblks || !d_offsets || nframes <= 0) return fail(J2K_ERR_INVALID_ARG, "bad buffers / nframes");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceCtx& d = ctx->devs[dev];
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : d.s_main;
    int kmax_max = 0;
    // the previous call's upload of the table may still be in flight from this vector
    CK(cudaStreamSynchronize(st));
    if ((rc = ht_block_kmax(s, cb_width, cb_height, kmax, d.he_kmax_host, kmax_max))) return rc;
    BlockTable* BT = nullptr;
    if ((rc = get_block_table(d, s, cb_width, cb_height, (long long)j2k_fwd_coeff_count(p), &BT))) return rc;
    if ((rc = d.he_kmax.ensure(d.he_kmax_host.size() + 16))) return rc;
    CK(cudaMemcpyAsync(d.he_kmax.p, d.he_kmax_host.data(), d.he_kmax_host.size(), cudaMemcpyHostToDevice, st));
    return launch_ht_encode(ctx, d, *BT, cb_width, cb_height, kmax_max, nframes, d_coeffs, (const unsigned char*)d.he_kmax.p, d_bytes,
                            (unsigned long long)bytes_cap, (HtBlock*)d_cblks, (unsigned long long*)d_offsets, st);
}

// Pixels up, compressed cleanup segments + records down.  Sub-batches on three streams with a lag of one: the size of
// sub-batch i's stream is read (a host wait on its kernels) after sub-batch i + 1 has been enqueued, then its bytes leave on
// the download stream while i + 1 computes.
int j2k_forward_ht(j2k_ctx* ctx, const j2k_fwd_params* p, int cb_width, int cb_height, int nframes, const void* pixels,
                   size_t frame_stride_bytes, const uint8_t* kmax, uint8_t* bytes_out, size_t bytes_cap, size_t* nbytes_out,
                   j2k_ht_cblk* cblks_out) {
    CtxGuard cg_(ctx);
    if (!ctx) return fail(J2K_ERR_INVALID_ARG, "context is NULL");
    if (!p) return fail(J2K_ERR_INVALID_ARG, "params is NULL");
    int rc = validate_ht_cb(cb_width, cb_height);
    if (rc) return rc;
    Spec s;
    if ((rc = spec_from_fwd(p, false, s))) return rc;
    if (!pixels || !kmax || !bytes_out || !cblks_out || !nbytes_out) return fail(J2K_ERR_INVALID_ARG, "NULL buffer");
    if (nframes <= 0) return fail(J2K_ERR_INVALID_ARG, "nframes must be positive");
    const size_t pixb = j2k_fwd_pixel_bytes(p);
    const long long cpf = (long long)j2k_fwd_coeff_count(p);
    const int bps = s.bit_depth <= 8 ? 1 : 2;
    if (frame_stride_bytes < pixb || frame_stride_bytes % bps) return fail(J2K_ERR_SIZE, "bad frame stride");
    std::lock_guard<std::mutex> lk(ctx->mu);
    // Frames are sharded over the context's devices in contiguous blocks like every host-batch call (no collective).  Each
    // device runs its own lagged pipeline; the host visits the devices round-robin, one sub-batch each per turn, and appends
    // a finished sub-batch's segments to the output as it collects it -- so with several devices the segments are NOT in
    // frame order in bytes_out; the records' offsets locate them.
    // bigger sub-batches than the transform-only paths (64 Msamples): the per-block packing kernel needs the blocks of several
    // frames to fill the GPU (measured on 32 C2 frames: 14.6 Gpixel/s at 16 Msamples, 20.9 at 64)
    const long long sub_samples = (long long)env_int("J2K_HT_SUBBATCH_MSAMPLES", 64) << 20;
    int sub = (int)(sub_samples / (cpf > 0 ? cpf : 1));
    if (sub < 1) sub = 1;
    struct Pend { int slot, f0, n; };
    struct DevRun {
        int di, next, f_end, it;
        bool have, used_out[2], used_k[2];
        Pend pend;
        BlockTable* BT;
    };
    const std::vector<int> H = healthy_devs(ctx);   // devices lost earlier stay out (j2k_device_failed)
    if (H.empty()) return fail(J2K_ERR_CUDA, "no usable device left in this context (every device failed)");
    const int nd = (int)H.size();
    const int per = (nframes + nd - 1) / nd;
    std::vector<DevRun> runs;
    int kmax_max = 0;
    size_t nblk = 0;
    for (int hk = 0; hk < nd; hk++) {
        const int di = H[hk];
        const int f0 = hk * per, f1 = f0 + per > nframes ? nframes : f0 + per;
        if (f0 >= f1) break;
        if ((rc = sync_dev(ctx, di))) return rc;   // the slot buffers may still serve an asynchronous job
        DeviceCtx& d = ctx->devs[di];
        DevRun r{};
        r.di = di; r.next = f0; r.f_end = f1;
        if ((rc = ht_block_kmax(s, cb_width, cb_height, kmax, d.he_kmax_host, kmax_max))) return rc;
        if ((rc = get_block_table(d, s, cb_width, cb_height, cpf, &r.BT))) return rc;
        nblk = (size_t)r.BT->nblocks;
        if ((rc = d.he_kmax.ensure(d.he_kmax_host.size() + 16))) return rc;
        CK(cudaMemcpyAsync(d.he_kmax.p, d.he_kmax_host.data(), d.he_kmax_host.size(), cudaMemcpyHostToDevice, d.s_main));
        runs.push_back(r);
    }
    size_t base = 0;
    auto finish = [&](DevRun& r, const Pend& q) -> int {
        int rc2 = set_dev(ctx, r.di);
        if (rc2) return rc2;
        DeviceCtx& d = ctx->devs[r.di];
        CK(cudaEventSynchronize(d.ev_k[q.slot]));
        const size_t count = (size_t)q.n * nblk;
        const unsigned long long tot = *(const unsigned long long*)d.he_pin[q.slot];
        const HtBlock* rec = (const HtBlock*)((const char*)d.he_pin[q.slot] + 16);
        if (base + tot > bytes_cap) {
            *nbytes_out = base + (size_t)tot;
            return fail(J2K_ERR_SIZE, "the compressed stream needs more than the %zu bytes provided (%zu so far)", bytes_cap, base + (size_t)tot);
        }
        if (tot) CK(cudaMemcpyAsync(bytes_out + base, d.he_bytes[q.slot].p, (size_t)tot, cudaMemcpyDeviceToHost, d.s_d2h));
        CK(cudaEventRecord(d.ev_out[q.slot], d.s_d2h));
        r.used_out[q.slot] = true;
        j2k_ht_cblk* dst = cblks_out + (size_t)q.f0 * nblk;
        for (size_t i = 0; i < count; i++) {
            dst[i].offset = rec[i].offset + base; dst[i].length = rec[i].length; dst[i].kmax = rec[i].kmax; dst[i].missing_msbs = rec[i].mmsb;
            dst[i].reserved = 0;
        }
        base += (size_t)tot;
        return 0;
    };
    auto enqueue = [&](DevRun& r) -> int {
        int rc2 = set_dev(ctx, r.di);
        if (rc2) return rc2;
        DeviceCtx& d = ctx->devs[r.di];
        const int b = r.next;
        const int nb = b + sub <= r.f_end ? sub : r.f_end - b;
        const int slot = r.it & 1;
        const size_t count = (size_t)nb * nblk;
        const size_t bound = ht_enc_bound(cb_width, cb_height, kmax_max, (long long)count);
        if ((rc2 = d.in[slot].ensure((size_t)nb * pixb))) return rc2;
        if ((rc2 = d.out[slot].ensure((size_t)nb * cpf * 4))) return rc2;
        if ((rc2 = d.he_bytes[slot].ensure(bound + 64))) return rc2;
        if ((rc2 = d.he_recs[slot].ensure(count * sizeof(HtBlock) + 64))) return rc2;
        if ((rc2 = d.he_off.ensure((count + 1) * 8 + 64))) return rc2;
        const size_t pin_need = 16 + count * sizeof(HtBlock);
        if (d.he_pin_cap[slot] < pin_need) {
            if (d.he_pin[slot]) cudaFreeHost(d.he_pin[slot]);
            d.he_pin[slot] = nullptr; d.he_pin_cap[slot] = 0;
            CK(cudaHostAlloc(&d.he_pin[slot], pin_need + pin_need / 4, cudaHostAllocPortable));
            d.he_pin_cap[slot] = pin_need + pin_need / 4;
        }
        if (r.used_k[slot]) CK(cudaStreamWaitEvent(d.s_h2d, d.ev_k[slot], 0));   // the slot's previous kernels have consumed `in`
        const unsigned char* src = (const unsigned char*)pixels + (size_t)b * frame_stride_bytes;
        if (frame_stride_bytes == pixb) CK(cudaMemcpyAsync(d.in[slot].p, src, (size_t)nb * pixb, cudaMemcpyHostToDevice, d.s_h2d));
        else CK(cudaMemcpy2DAsync(d.in[slot].p, pixb, src, frame_stride_bytes, pixb, nb, cudaMemcpyHostToDevice, d.s_h2d));
        CK(cudaEventRecord(d.ev_in[slot], d.s_h2d));
        CK(cudaStreamWaitEvent(d.s_main, d.ev_in[slot], 0));
        if (r.used_out[slot]) CK(cudaStreamWaitEvent(d.s_main, d.ev_out[slot], 0));   // its previous stream has left the device
        Plan* P = nullptr;
        if ((rc2 = get_plan(d, s, p, sizeof *p, nb, (long long)(pixb / bps), &P))) return rc2;
        rc2 = run_plan(ctx, *P, d.in[slot].p, d.out[slot].p, nullptr, false, d.s_main);
        if (rc2 < 0) return rc2;
        if ((rc2 = launch_ht_encode(ctx, d, *r.BT, cb_width, cb_height, kmax_max, nb, (const int32_t*)d.out[slot].p,
                                    (const unsigned char*)d.he_kmax.p, (unsigned char*)d.he_bytes[slot].p, (unsigned long long)bound,
                                    (HtBlock*)d.he_recs[slot].p, (unsigned long long*)d.he_off.p, d.s_main))) return rc2;
        CK(cudaMemcpyAsync(d.he_pin[slot], (const unsigned long long*)d.he_off.p + count, 8, cudaMemcpyDeviceToHost, d.s_main));
        CK(cudaMemcpyAsync((char*)d.he_pin[slot] + 16, d.he_recs[slot].p, count * sizeof(HtBlock), cudaMemcpyDeviceToHost, d.s_main));
        CK(cudaEventRecord(d.ev_k[slot], d.s_main));
        r.used_k[slot] = true;
        const Pend now{slot, b, nb};
        r.next = b + nb; r.it++;
        // the size of the previous sub-batch's stream is read (a host wait on its kernels) now that this one is enqueued;
        // its bytes then leave on the download stream while this one computes
        if (r.have && (rc2 = finish(r, r.pend))) return rc2;
        r.pend = now; r.have = true;
        return 0;
    };
    for (bool more = true; more && rc == 0;) {
        more = false;
        for (DevRun& r : runs) {
            if (rc) break;
            if (r.next < r.f_end) rc = enqueue(r);
            else if (r.have) { rc = finish(r, r.pend); r.have = false; }
            if (r.next < r.f_end || r.have) more = true;
        }
    }
    for (DevRun& r : runs) {
        int r2 = sync_dev(ctx, r.di);
        if (rc == 0) rc = r2;
        DeviceCtx& d = ctx->devs[r.di];
        d.ev_k_used[0] = d.ev_k_used[1] = false; d.ev_out_used[0] = d.ev_out_used[1] = false;
    }
    if (rc == 0) *nbytes_out = base;
    return rc < 0 ? rc : J2K_OK;
}

// ---- asynchronous

// A ticket completes when the job's last device-to-host copy has landed: one event per device the job used, recorded
// behind that copy, so waiting for ticket i does not wait for the jobs submitted after it (frame i+1 can be in
// flight while the caller entropy-codes frame i).
static int64_t make_ticket(j2k_ctx* ctx, const std::vector<int>& used) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    std::vector<j2k_ctx::TicketEv> evs;
    for (int di : used) {
        DeviceCtx& d = ctx->devs[di];
        cudaEvent_t ev = nullptr;
        if (cudaSetDevice(d.dev) != cudaSuccess || cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventRecord(ev, d.s_d2h) != cudaSuccess) {
            for (auto& t : evs) cudaEventDestroy(t.ev);
            if (ev) cudaEventDestroy(ev);
            return fail(J2K_ERR_CUDA, "ticket event: %s", cudaGetErrorString(cudaGetLastError()));
        }
        evs.push_back({di, ev});
    }
    long long t = ctx->next_ticket++;
    ctx->tickets[t] = evs;
    return t;
}

int64_t j2k_submit_forward(j2k_ctx* ctx, const j2k_fwd_params* p, int nframes, const void* pixels, size_t frame_stride_bytes,
                           int32_t* coeffs_out) {
    CtxGuard cg_(ctx);
    std::vector<int> used;
    int rc = forward_host(ctx, p, nframes, pixels, frame_stride_bytes, coeffs_out, false, false, &used);
    if (rc) return rc;
    return make_ticket(ctx, used);
}

int64_t j2k_submit_inverse(j2k_ctx* ctx, const j2k_inv_params* p, int nframes, const int32_t* coeffs_in, void* pixels_out,
                           size_t frame_stride_bytes, int32_t* planes_out) {
    CtxGuard cg_(ctx);
    std::vector<int> used;
    int rc = inverse_host(ctx, p, nframes, coeffs_in, pixels_out, frame_stride_bytes, planes_out, false, &used);
    if (rc) return rc;
    return make_ticket(ctx, used);
}

int j2k_wait(j2k_ctx* ctx, int64_t ticket) {
    CtxGuard cg_(ctx);
    if (!ctx) return fail(J2K_ERR_INVALID_ARG, "context is NULL");
    std::vector<j2k_ctx::TicketEv> evs;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        auto it = ctx->tickets.find(ticket);
        if (it == ctx->tickets.end()) return fail(J2K_ERR_TICKET, "unknown ticket %lld", (long long)ticket);
        evs = it->second;
        ctx->tickets.erase(it);
    }
    int rc = 0;
    for (auto& t : evs) {
        cudaError_t e = cudaSetDevice(ctx->devs[t.di].dev);
        if (e == cudaSuccess) e = cudaEventSynchronize(t.ev);
        cudaEventDestroy(t.ev);
        if (e != cudaSuccess && !rc) rc = fail(J2K_ERR_CUDA, "j2k_wait: %s", cudaGetErrorString(e));
    }
    return rc;
}

// ---- wavelet package API: in place on a host plane, origin (x0, y0), stride = width

static int dwt_api(j2k_ctx* ctx, void* data, int width, int height, int levels, int x0, int y0, bool reversible, bool fwd, bool f64 = false) {
    CtxGuard cg_(ctx);
    if (!ctx || !data) return fail(J2K_ERR_INVALID_ARG, "NULL argument");
    if (width <= 0 || height <= 0) return fail(J2K_ERR_INVALID_ARG, "invalid dimensions: %dx%d", width, height);
    if (levels < 0) levels = 0;
    if (levels > 30) levels = 30;  // windows are 1x1 long before that
    if (x0 < 0 || y0 < 0) return fail(J2K_ERR_INVALID_ARG, "negative origin");
    int rc = set_dev(ctx, 0);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceCtx& d = ctx->devs[0];
    Spec s{};
    s.fwd = fwd; s.W = width; s.H = height; s.C = 1; s.bit_depth = 16; s.is_signed = 1; s.L = levels; s.reversible = reversible;
    s.mct_mode = J2K_MCT_NONE; s.direct = true;
    TileGeom g{0, 0, width, height, x0, y0, 0};
    s.tiles.push_back(g);
    struct { int w, h, l, x, y, r, f; } blob = {width, height, levels, x0, y0, reversible, fwd};
    Plan* P = nullptr;
    if ((rc = get_plan(d, s, &blob, sizeof blob, 1, (long long)width * height, &P))) return rc;
    size_t bytes = (size_t)width * height * 4;
    if ((rc = d.api[0].ensure(bytes)) || (rc = d.api[1].ensure(bytes))) return rc;
    const long long npx = (long long)width * height;
    if (f64) {  // the float64 wrappers (dwt97.go:345-351,415-421): ConvertFloat64ToFloat32, the float32 transform, and back
        if ((rc = d.api[2].ensure(bytes * 2))) return rc;
        CK(cudaMemcpyAsync(d.api[2].p, data, bytes * 2, cudaMemcpyHostToDevice, d.s_main));
        J2K_LAUNCH(f64_to_f32_kernel, (unsigned)((npx + 255) / 256), 256, d.s_main, (const double*)d.api[2].p, (float*)d.api[0].p, npx);
        CK(cudaGetLastError());
        ctx->launches++;
    } else
    CK(cudaMemcpyAsync(d.api[0].p, data, bytes, cudaMemcpyHostToDevice, d.s_main));
    // the untouched part of the plane (nothing, when at least one level runs) keeps the input values
    CK(cudaMemcpyAsync(d.api[1].p, d.api[0].p, bytes, cudaMemcpyDeviceToDevice, d.s_main));
    if (fwd) rc = run_plan(ctx, *P, d.api[0].p, d.api[1].p, nullptr, false, d.s_main);
    else rc = run_plan(ctx, *P, nullptr, d.api[0].p, d.api[1].p, false, d.s_main);
    if (rc < 0) return rc;
    if (f64) {
        J2K_LAUNCH(f32_to_f64_kernel, (unsigned)((npx + 255) / 256), 256, d.s_main, (const float*)d.api[1].p, (double*)d.api[2].p, npx);
        CK(cudaGetLastError());
        ctx->launches++;
        CK(cudaMemcpyAsync(data, d.api[2].p, bytes * 2, cudaMemcpyDeviceToHost, d.s_main));
    } else
    CK(cudaMemcpyAsync(data, d.api[1].p, bytes, cudaMemcpyDeviceToHost, d.s_main));
    CK(cudaStreamSynchronize(d.s_main));
    return J2K_OK;
}

int j2k_dwt97_forward_f64(j2k_ctx* ctx, double* data, int w, int h, int levels, int x0, int y0) { return dwt_api(ctx, data, w, h, levels, x0, y0, false, true, true); }
int j2k_dwt97_inverse_f64(j2k_ctx* ctx, double* data, int w, int h, int levels, int x0, int y0) { return dwt_api(ctx, data, w, h, levels, x0, y0, false, false, true); }

// wavelet.LLDimensionsWithParity (layout.go:11-33); pure host arithmetic
int j2k_ll_dimensions(int width, int height, int levels, int x0, int y0, int* ll_width, int* ll_height) {
    if (!ll_width || !ll_height) return fail(J2K_ERR_INVALID_ARG, "NULL argument");
    if (width <= 0 || height <= 0) { *ll_width = 0; *ll_height = 0; return J2K_OK; }
    int cw = width, ch = height, cx = x0, cy = y0;
    for (int l = 0; l < levels; l++) {
        if (cw <= 1 && ch <= 1) break;
        cw = split_len(cw, (cx & 1) == 0); ch = split_len(ch, (cy & 1) == 0);   // nextLowpassWindow, layout.go:35-44
        cx = next_coord(cx); cy = next_coord(cy);
    }
    *ll_width = cw; *ll_height = ch;
    return J2K_OK;
}

int j2k_convert_f64_to_i32(j2k_ctx* ctx, const double* in, int32_t* out, size_t n) {
    CtxGuard cg_(ctx);
    if (!ctx) return fail(J2K_ERR_INVALID_ARG, "context is NULL");
    if (n == 0) return J2K_OK;
    if (!in || !out) return fail(J2K_ERR_INVALID_ARG, "NULL buffer");
    int rc = set_dev(ctx, 0);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceCtx& d = ctx->devs[0];
    if ((rc = d.api[0].ensure(n * 8)) || (rc = d.api[1].ensure(n * 4))) return rc;
    CK(cudaMemcpyAsync(d.api[0].p, in, n * 8, cudaMemcpyHostToDevice, d.s_main));
    J2K_LAUNCH(f64_to_i32_kernel, (unsigned)((n + 255) / 256), 256, d.s_main, (const double*)d.api[0].p, (int*)d.api[1].p, (long long)n);
    CK(cudaGetLastError());
    ctx->launches++;
    CK(cudaMemcpyAsync(out, d.api[1].p, n * 4, cudaMemcpyDeviceToHost, d.s_main));
    CK(cudaStreamSynchronize(d.s_main));
    return J2K_OK;
}

// colorspace.InterleaveComponents / DeinterleaveComponents (rgb.go:54-98)
static int interleave_api(j2k_ctx* ctx, const int32_t* const* planes_in, int32_t* const* planes_out, int32_t* inter_out, const int32_t* inter_in,
                          size_t n, int C) {
    CtxGuard cg_(ctx);
    if (!ctx) return fail(J2K_ERR_INVALID_ARG, "context is NULL");
    if (C <= 0 || C > 16) return fail(J2K_ERR_INVALID_ARG, "invalid number of components: %d", C);
    if (n == 0) return J2K_OK;
    int rc = set_dev(ctx, 0);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceCtx& d = ctx->devs[0];
    const size_t bytes = n * (size_t)C * 4;
    if ((rc = d.api[0].ensure(bytes)) || (rc = d.api[1].ensure(bytes))) return rc;
    const bool to_inter = planes_in != nullptr;
    if (to_inter) {
        for (int c = 0; c < C; c++) {
            if (!planes_in[c]) return fail(J2K_ERR_INVALID_ARG, "component %d is NULL", c);
            CK(cudaMemcpyAsync((char*)d.api[0].p + (size_t)c * n * 4, planes_in[c], n * 4, cudaMemcpyHostToDevice, d.s_main));
        }
    } else {
        CK(cudaMemcpyAsync(d.api[0].p, inter_in, bytes, cudaMemcpyHostToDevice, d.s_main));
    }
    J2K_LAUNCH(interleave_kernel, (unsigned)((n * C + 255) / 256), 256, d.s_main, (const int*)d.api[0].p, (int*)d.api[1].p, (long long)n, C, to_inter ? 1 : 0);
    CK(cudaGetLastError());
    ctx->launches++;
    if (to_inter) CK(cudaMemcpyAsync(inter_out, d.api[1].p, bytes, cudaMemcpyDeviceToHost, d.s_main));
    else
        for (int c = 0; c < C; c++) CK(cudaMemcpyAsync(planes_out[c], (char*)d.api[1].p + (size_t)c * n * 4, n * 4, cudaMemcpyDeviceToHost, d.s_main));
    CK(cudaStreamSynchronize(d.s_main));
    return J2K_OK;
}

int j2k_interleave_components(j2k_ctx* ctx, const int32_t* const* components, int n_components, size_t n_pixels, int32_t* out) {
    CtxGuard cg_(ctx);
    if (!components || !out) return fail(J2K_ERR_INVALID_ARG, "NULL buffer");
    return interleave_api(ctx, components, nullptr, out, nullptr, n_pixels, n_components);
}

int j2k_deinterleave_components(j2k_ctx* ctx, const int32_t* data, size_t n_pixels, int n_components, int32_t* const* components_out) {
    CtxGuard cg_(ctx);
    if (!data || !components_out) return fail(J2K_ERR_INVALID_ARG, "NULL buffer");
    for (int c = 0; c < n_components; c++)
        if (!components_out[c]) return fail(J2K_ERR_INVALID_ARG, "component %d is NULL", c);
    return interleave_api(ctx, nullptr, components_out, nullptr, data, n_pixels, n_components);
}

int j2k_dwt53_forward(j2k_ctx* ctx, int32_t* data, int w, int h, int levels, int x0, int y0) { return dwt_api(ctx, data, w, h, levels, x0, y0, true, true); }
int j2k_dwt53_inverse(j2k_ctx* ctx, int32_t* data, int w, int h, int levels, int x0, int y0) { return dwt_api(ctx, data, w, h, levels, x0, y0, true, false); }
int j2k_dwt97_forward(j2k_ctx* ctx, float* data, int w, int h, int levels, int x0, int y0) { return dwt_api(ctx, data, w, h, levels, x0, y0, false, true); }
int j2k_dwt97_inverse(j2k_ctx* ctx, float* data, int w, int h, int levels, int x0, int y0) { return dwt_api(ctx, data, w, h, levels, x0, y0, false, false); }

// ---- pointwise package APIs

static int api3(j2k_ctx* ctx, int op, size_t n, const int32_t* a, const int32_t* b, const int32_t* c, int32_t* x, int32_t* y, int32_t* z) {
    CtxGuard cg_(ctx);
    if (!ctx) return fail(J2K_ERR_INVALID_ARG, "context is NULL");
    if (n == 0) return J2K_OK;
    if (!a || !b || !c || !x || !y || !z) return fail(J2K_ERR_INVALID_ARG, "NULL buffer");
    int rc = set_dev(ctx, 0);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceCtx& d = ctx->devs[0];
    const int32_t* hin[3] = {a, b, c};
    int32_t* hout[3] = {x, y, z};
    for (int i = 0; i < 6; i++) if ((rc = d.api[i].ensure(n * 4))) return rc;
    for (int i = 0; i < 3; i++) CK(cudaMemcpyAsync(d.api[i].p, hin[i], n * 4, cudaMemcpyHostToDevice, d.s_main));
    J2K_LAUNCH(color_api_kernel, (unsigned)((n + 255) / 256), 256, d.s_main, op, (long long)n, (const int*)d.api[0].p, (const int*)d.api[1].p,
               (const int*)d.api[2].p, (int*)d.api[3].p, (int*)d.api[4].p, (int*)d.api[5].p);
    CK(cudaGetLastError());
    ctx->launches++;
    for (int i = 0; i < 3; i++) CK(cudaMemcpyAsync(hout[i], d.api[3 + i].p, n * 4, cudaMemcpyDeviceToHost, d.s_main));
    CK(cudaStreamSynchronize(d.s_main));
    return J2K_OK;
}

int j2k_rct_forward(j2k_ctx* ctx, size_t n, const int32_t* r, const int32_t* g, const int32_t* b, int32_t* y, int32_t* cb, int32_t* cr) { return api3(ctx, 0, n, r, g, b, y, cb, cr); }
int j2k_rct_inverse(j2k_ctx* ctx, size_t n, const int32_t* y, const int32_t* cb, const int32_t* cr, int32_t* r, int32_t* g, int32_t* b) { return api3(ctx, 1, n, y, cb, cr, r, g, b); }
int j2k_ict_forward(j2k_ctx* ctx, size_t n, const int32_t* r, const int32_t* g, const int32_t* b, int32_t* y, int32_t* cb, int32_t* cr) { return api3(ctx, 2, n, r, g, b, y, cb, cr); }
int j2k_ict_inverse(j2k_ctx* ctx, size_t n, const int32_t* y, const int32_t* cb, const int32_t* cr, int32_t* r, int32_t* g, int32_t* b) { return api3(ctx, 3, n, y, cb, cr, r, g, b); }

// colorspace.ConvertRGBToYCbCr / ConvertYCbCrToRGB (rgb.go:17-52): the ICT on an interleaved RGB image
int j2k_rgb_to_ycbcr(j2k_ctx* ctx, const int32_t* rgb, int width, int height, int32_t* y, int32_t* cb, int32_t* cr) {
    CtxGuard cg_(ctx);
    if (!ctx) return fail(J2K_ERR_INVALID_ARG, "context is NULL");
    if (width <= 0 || height <= 0) return J2K_OK;
    if (!rgb || !y || !cb || !cr) return fail(J2K_ERR_INVALID_ARG, "NULL buffer");
    const size_t n = (size_t)width * height;
    int rc = set_dev(ctx, 0);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceCtx& d = ctx->devs[0];
    if ((rc = d.api[0].ensure(n * 12)) || (rc = d.api[1].ensure(n * 12)) || (rc = d.api[2].ensure(n * 12))) return rc;
    CK(cudaMemcpyAsync(d.api[0].p, rgb, n * 12, cudaMemcpyHostToDevice, d.s_main));
    J2K_LAUNCH(interleave_kernel, (unsigned)((n * 3 + 255) / 256), 256, d.s_main, (const int*)d.api[0].p, (int*)d.api[1].p, (long long)n, 3, 0);
    const int* pl = (const int*)d.api[1].p;
    int* o = (int*)d.api[2].p;
    J2K_LAUNCH(color_api_kernel, (unsigned)((n + 255) / 256), 256, d.s_main, 2, (long long)n, pl, pl + n, pl + 2 * n, o, o + n, o + 2 * n);
    CK(cudaGetLastError());
    ctx->launches += 2;
    int32_t* hout[3] = {y, cb, cr};
    for (int i = 0; i < 3; i++) CK(cudaMemcpyAsync(hout[i], o + (size_t)i * n, n * 4, cudaMemcpyDeviceToHost, d.s_main));
    CK(cudaStreamSynchronize(d.s_main));
    return J2K_OK;
}

int j2k_ycbcr_to_rgb(j2k_ctx* ctx, const int32_t* y, const int32_t* cb, const int32_t* cr, int width, int height, int32_t* rgb) {
    CtxGuard cg_(ctx);
    if (!ctx) return fail(J2K_ERR_INVALID_ARG, "context is NULL");
    if (width <= 0 || height <= 0) return J2K_OK;
    if (!rgb || !y || !cb || !cr) return fail(J2K_ERR_INVALID_ARG, "NULL buffer");
    const size_t n = (size_t)width * height;
    int rc = set_dev(ctx, 0);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceCtx& d = ctx->devs[0];
    if ((rc = d.api[0].ensure(n * 12)) || (rc = d.api[1].ensure(n * 12)) || (rc = d.api[2].ensure(n * 12))) return rc;
    const int32_t* hin[3] = {y, cb, cr};
    int* pl = (int*)d.api[0].p;
    for (int i = 0; i < 3; i++) CK(cudaMemcpyAsync(pl + (size_t)i * n, hin[i], n * 4, cudaMemcpyHostToDevice, d.s_main));
    int* o = (int*)d.api[1].p;
    J2K_LAUNCH(color_api_kernel, (unsigned)((n + 255) / 256), 256, d.s_main, 3, (long long)n, (const int*)pl, (const int*)(pl + n), (const int*)(pl + 2 * n),
               o, o + n, o + 2 * n);
    J2K_LAUNCH(interleave_kernel, (unsigned)((n * 3 + 255) / 256), 256, d.s_main, (const int*)o, (int*)d.api[2].p, (long long)n, 3, 1);
    CK(cudaGetLastError());
    ctx->launches += 2;
    CK(cudaMemcpyAsync(rgb, d.api[2].p, n * 12, cudaMemcpyDeviceToHost, d.s_main));
    CK(cudaStreamSynchronize(d.s_main));
    return J2K_OK;
}

static int api1(j2k_ctx* ctx, int op, const void* in, void* out, size_t n, double step) {
    CtxGuard cg_(ctx);
    if (!ctx) return fail(J2K_ERR_INVALID_ARG, "context is NULL");
    if (n == 0) return J2K_OK;
    if (!in || !out) return fail(J2K_ERR_INVALID_ARG, "NULL buffer");
    if (op < 2 && step <= 0) { memmove(out, in, n * 4); return J2K_OK; }  // quantization.go:311-314,327-330: returns the input
    int rc = set_dev(ctx, 0);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceCtx& d = ctx->devs[0];
    if ((rc = d.api[0].ensure(n * 4)) || (rc = d.api[1].ensure(n * 4))) return rc;
    CK(cudaMemcpyAsync(d.api[0].p, in, n * 4, cudaMemcpyHostToDevice, d.s_main));
    unsigned grid = (unsigned)((n + 255) / 256);
    if (op < 2) J2K_LAUNCH(quant_api_kernel, grid, 256, d.s_main, op, (long long)n, (const int*)d.api[0].p, (int*)d.api[1].p, step);
    else J2K_LAUNCH(f32_to_i32_kernel, grid, 256, d.s_main, (const float*)d.api[0].p, (int*)d.api[1].p, (long long)n);
    CK(cudaGetLastError());
    ctx->launches++;
    CK(cudaMemcpyAsync(out, d.api[1].p, n * 4, cudaMemcpyDeviceToHost, d.s_main));
    CK(cudaStreamSynchronize(d.s_main));
    return J2K_OK;
}

int j2k_quantize_coefficients(j2k_ctx* ctx, const int32_t* in, int32_t* out, size_t n, double step) { return api1(ctx, 0, in, out, n, step); }
int j2k_dequantize_coefficients(j2k_ctx* ctx, const int32_t* in, int32_t* out, size_t n, double step) { return api1(ctx, 1, in, out, n, step); }
int j2k_convert_f32_to_i32(j2k_ctx* ctx, const float* in, int32_t* out, size_t n) { return api1(ctx, 2, in, out, n, 1.0); }

// ---- quantization.go scalar metadata (host only)

static const double kNorms97[4][10] = {  // quantization.go:17-22
    {1.000, 1.965, 4.177, 8.403, 16.90, 33.84, 67.69, 135.3, 270.6, 540.9},
    {2.022, 3.989, 8.355, 17.04, 34.27, 68.63, 137.3, 274.6, 549.0, 0.0},
    {2.022, 3.989, 8.355, 17.04, 34.27, 68.63, 137.3, 274.6, 549.0, 0.0},
    {2.080, 3.865, 8.307, 17.18, 34.71, 69.59, 139.3, 278.6, 557.2, 0.0},
};

static double norm97(int level, int orient) {  // quantization.go:39-52
    if (level < 0) level = 0;
    if (orient == 0 && level >= 10) level = 9;
    else if (orient > 0 && level >= 9) level = 8;
    if (orient < 0 || orient > 3) return 1.0;
    return kNorms97[orient][level];
}

static void band_params(int idx, int L, int* orient, int* level) {  // quantization.go:68-83
    int resno = 0;
    if (idx == 0) *orient = 0;
    else { resno = (idx - 1) / 3 + 1; *orient = (idx - 1) % 3 + 1; }
    *level = L - resno;
    if (*level < 0) *level = 0;
}

static uint16_t encode_step(double step, int numbps) {  // quantization.go:102-128
    if (step <= 0) return 0;
    int32_t fixed = (int32_t)floor(step * 8192.0);
    if (fixed <= 0) fixed = 1;
    int lg = 0;
    for (uint32_t v = (uint32_t)fixed; v > 1; v >>= 1) lg++;
    int pw = lg - 13, n = 11 - lg;
    int32_t mant = n < 0 ? (fixed >> -n) : (int32_t)((uint32_t)fixed << n);
    mant &= 0x7ff;
    int expn = numbps - pw;
    if (expn < 0) expn = 0;
    if (expn > 0x1f) expn = 0x1f;
    return (uint16_t)((expn << 11) | mant);
}

int j2k_quant_openjpeg_params(int L, int bit_depth, uint16_t* encoded, double* step_sizes) {
    if (!encoded || !step_sizes) return fail(J2K_ERR_INVALID_ARG, "NULL argument");
    if (L < 0) L = 0;
    if (L > J2K_MAX_LEVELS) return fail(J2K_ERR_INVALID_ARG, "too many levels");
    int nb = 3 * L + 1;
    for (int b = 0; b < nb; b++) {
        int o, lv;
        band_params(b, L, &o, &lv);
        double norm = norm97(lv, o), st = 1.0;
        if (norm > 0) st = 1.0 / norm;
        step_sizes[b] = st;
        encoded[b] = encode_step(st, bit_depth);
    }
    return nb;
}

int j2k_quant_quality_params(int quality, int L, int bit_depth, uint16_t* encoded, double* step_sizes) {
    if (!encoded || !step_sizes) return fail(J2K_ERR_INVALID_ARG, "NULL argument");
    if (L > J2K_MAX_LEVELS) return fail(J2K_ERR_INVALID_ARG, "too many levels");
    if (quality < 1) quality = 1;
    if (quality > 100) quality = 100;
    double scale = pow(2.0, (100.0 - (double)quality) / 12.5);  // quantization.go:54-66
    if (scale < 0.01) scale = 0.01;
    scale *= 0.05;
    int nb;
    if (L <= 0) { nb = 1; step_sizes[0] = scale; }
    else {
        nb = 3 * L + 1;
        for (int b = 0; b < nb; b++) {
            int o, lv;
            band_params(b, L, &o, &lv);
            double norm = norm97(lv, o);
            step_sizes[b] = norm <= 0 ? scale : scale / norm;
        }
    }
    for (int b = 0; b < nb; b++) encoded[b] = encode_step(step_sizes[b], bit_depth);
    return nb;
}

int j2k_quant_runtime_steps(const uint16_t* encoded, int n, int L, int bit_depth, double* steps) {
    if (!encoded || !steps || n < 0) return fail(J2K_ERR_INVALID_ARG, "bad argument");
    for (int i = 0; i < n; i++) {
        int o, lv;
        band_params(i, L, &o, &lv);
        int gain = o == 3 ? 2 : (o == 1 || o == 2) ? 1 : 0;
        int expn = (encoded[i] >> 11) & 0x1f;
        double mant = (double)(encoded[i] & 0x7ff);
        steps[i] = (double)(float)ldexp(1.0 + mant / 2048.0, bit_depth + gain - expn);  // quantization.go:130-135,151
    }
    return n;
}

int j2k_quant_decode_steps(const uint16_t* encoded, int n, int L, int bit_depth, int reversible, double* steps) {
    (void)L;
    if (!encoded || !steps || n < 0) return fail(J2K_ERR_INVALID_ARG, "bad argument");
    for (int i = 0; i < n; i++) {
        int expn = (encoded[i] >> 11) & 0x1f, mant = encoded[i] & 0x7ff;
        int gain = 0;
        if (reversible && i != 0) gain = ((i - 1) % 3 + 1) == 3 ? 2 : 1;  // t2/tile_decoder.go:1048-1060
        steps[i] = ldexp(1.0 + (double)mant / 2048.0, bit_depth + gain - expn);  // t2/tile_decoder.go:1062-1065
    }
    return n;
}

}  // extern "C"
#pragma GCC visibility pop
