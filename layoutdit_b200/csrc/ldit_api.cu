// extern "C" boundary of libldit_b200.so -- see include/ldit.h for the contract.
// Host code here only validates arguments, builds TMA tensor maps and enqueues kernels.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <utility>
#include <vector>

#include <map>
#include <mutex>
#include <unordered_map>

#include "../../include/ldit.h"
#include "attention_v3.cuh"
#include "backward.cuh"
#include "attention_bwd_tc.cuh"
#include "gemm.cuh"
#include "rowwise.cuh"
// Superseded / experimental kernels (round-1 attention variants, the fused fc1+fc2 kernel): measured slower than the
// product path, kept for A/B only.  They are compiled in with -DLDIT_EXPERIMENTAL; without it their entry points
// return LDIT_E_UNSUPPORTED.
#ifdef LDIT_EXPERIMENTAL
#include "attention_mma.cuh"
#include "attention_tc.cuh"
#include "attention_tc2.cuh"
#include "mlp_fused.cuh"
#endif

using namespace ldit;

namespace {

std::atomic<unsigned long long> g_launches{0};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    (void)cudaGetLastError();
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// Encoded tensor maps are a pure function of (pointer, dtype, dims, strides, box, swizzle): the workspaces and weights
// of a forward keep their addresses from call to call, so the ~1 us driver call per map (3 per GEMM, 140 per forward)
// is paid once.  Bounded; cleared wholesale when full (a stale entry can only be hit by an identical request, for
// which it is still correct).
struct TmapKey {
  uint64_t w[8];
  bool operator==(const TmapKey& o) const { return memcmp(w, o.w, sizeof(w)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t v : k.w) { h ^= v; h *= 1099511628211ull; }
    return static_cast<size_t>(h);
  }
};
std::mutex g_tmap_mu;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;

int encode_cached(CUtensorMap* tm, CUtensorMapDataType dt, uint32_t rank, const void* ptr, const cuuint64_t* dims,
                  const cuuint64_t* strides, const cuuint32_t* box, CUtensorMapSwizzle swz) {
  TmapKey key{};
  key.w[0] = reinterpret_cast<uint64_t>(ptr);
  key.w[1] = (static_cast<uint64_t>(dt) << 40) | (static_cast<uint64_t>(rank) << 32) | static_cast<uint64_t>(swz);
  for (uint32_t i = 0; i < rank; ++i) {
    key.w[2] = key.w[2] * 0x9E3779B97F4A7C15ull + dims[i];
    key.w[3] = key.w[3] * 0x9E3779B97F4A7C15ull + (i + 1 < rank ? strides[i] : 0);
    key.w[4] = key.w[4] * 1000003ull + box[i];
    key.w[5 + (i % 3)] ^= dims[i] << (13 * (i / 3)) ^ (static_cast<uint64_t>(box[i]) << 40);
  }
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) { *tm = it->second; return 0; }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return LDIT_E_NO_DRIVER;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(tm, dt, rank, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return 10000 + static_cast<int>(r);
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  if (g_tmap_cache.size() >= 8192) g_tmap_cache.clear();
  g_tmap_cache.emplace(key, *tm);
  return 0;
}

// 2-D row-major tensor [rows, cols], box [box_rows, box_cols], out-of-bounds elements read as
// zero (loads) / are clipped (stores).
int make_tmap_2d(CUtensorMap* tm, const void* ptr, CUtensorMapDataType dt, int esize, uint64_t rows, uint64_t cols,
                 uint64_t pitch_elems, uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swz) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_elems * esize};
  cuuint32_t box[2] = {box_cols, box_rows};
  return encode_cached(tm, dt, 2, ptr, dims, strides, box, swz);
}

// bf16 GEMM operand: box [box_rows, 64 cols] = 128-byte rows, 128-byte swizzle
int make_tmap_bf16_2d(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  return make_tmap_2d(tm, ptr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, rows, cols, cols, box_rows, 64, CU_TENSOR_MAP_SWIZZLE_128B);
}

// qkv viewed as [B, N, 3D] bf16: box [1, box_rows, 64 cols], 128-byte swizzle.  The image is its
// own dimension so that rows past the end of an image read as zero instead of the next image.
int make_tmap_qkv_3d(CUtensorMap* tm, const void* ptr, uint64_t B, uint64_t N, uint64_t cols, uint32_t box_rows);
int make_tmap_rows_3d(CUtensorMap* tm, const void* ptr, uint64_t B, uint64_t N, uint64_t cols, uint32_t box_rows) {
  return make_tmap_qkv_3d(tm, ptr, B, N, cols, box_rows);
}
int make_tmap_qkv_3d(CUtensorMap* tm, const void* ptr, uint64_t B, uint64_t N, uint64_t cols, uint32_t box_rows) {
  cuuint64_t dims[3] = {cols, N, B};
  cuuint64_t strides[2] = {cols * 2, N * cols * 2};
  cuuint32_t box[3] = {64, box_rows, 1};
  return encode_cached(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ptr, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

// channels-last image [B, H, W, C] bf16 as a 4-D tensor (C fastest): box [1, box_h, box_w, box_c]
int make_tmap_nhwc_4d(CUtensorMap* tm, const void* ptr, uint64_t B, uint64_t H, uint64_t W, uint64_t C, uint32_t box_c,
                      uint32_t box_w, uint32_t box_h, CUtensorMapSwizzle swz) {
  cuuint64_t dims[4] = {C, W, H, B};
  cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
  cuuint32_t box[4] = {box_c, box_w, box_h, 1};
  return encode_cached(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, dims, strides, box, swz);
}

#ifdef LDIT_DEBUG_HOOKS
long long* g_gemm_tl = nullptr;   // diagnosis builds only: device buffer for the GEMM timeline
#endif
long long* g_attn_dbg = nullptr;  // -DLDIT_A3_TIMELINE builds only: device buffer for the attention timeline
std::atomic<int> g_attn_impl{-1};  // 0 = attention_v3 (default); 1, 2, 4 = the superseded variants (-DLDIT_EXPERIMENTAL)

// Per-device host state (function attributes, SM count and the persisting-L2 limit are per device / context; the
// library may be driven for several devices from one process).
struct DeviceState {
  int sms = 0;
  size_t persist_limit = 0;
  std::map<const void*, size_t> smem_attr;   // kernel -> largest dynamic smem size set so far
};
std::mutex g_dev_mu;
std::map<int, DeviceState> g_dev;

DeviceState& dev_state_locked(int* dev_out = nullptr) {   // caller holds g_dev_mu
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); dev = 0; }
  if (dev_out) *dev_out = dev;
  DeviceState& d = g_dev[dev];
  if (d.sms == 0) {
    int n = 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { (void)cudaGetLastError(); n = 148; }
    d.sms = n > 0 ? n : 148;
  }
  return d;
}

int num_sms() {
  std::lock_guard<std::mutex> lk(g_dev_mu);
  return dev_state_locked().sms;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize (and, optionally, the largest shared-memory carve-out) for `kernel` on the
// current device, once per (device, kernel, size growth)
template <typename K>
cudaError_t ensure_smem(K kernel, size_t bytes, bool max_carveout = false) {
  std::lock_guard<std::mutex> lk(g_dev_mu);
  DeviceState& d = dev_state_locked();
  size_t& have = d.smem_attr[reinterpret_cast<const void*>(kernel)];
  if (bytes <= have) return cudaSuccess;
  cudaError_t e = cudaSuccess;
  if (bytes > 48 * 1024) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e == cudaSuccess && max_carveout)
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return e; }
  have = bytes;
  return cudaSuccess;
}

std::atomic<int> g_pdl{-1};  // programmatic dependent launch: -1 = read LDIT_PDL once; 0 off; 1 on (default)
bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("LDIT_PDL");
    v = (e && atoi(e) == 0) ? 0 : 1;
    g_pdl.store(v);
  }
  return v != 0;
}

// Optional L2 persistence window (ldit_set_l2_window): every launch ON THAT STREAM carries an access-policy window
// over the fp32 residual stream, the one buffer the whole forward keeps coming back to (LayerNorm reads, TMA
// reduce-adds, taps) while 58-77 MB activations stream through the same L2 between two visits.  Keyed by stream, so
// two engines (or two devices) driven from one process do not see each other's window.
struct L2Window { void* ptr; size_t bytes; float ratio; };
std::mutex g_win_mu;
std::map<cudaStream_t, L2Window> g_windows;
std::atomic<int> g_num_windows{0};

// Every kernel goes through here: cluster dimension + programmatic stream serialization (the
// kernel may start its prologue while its predecessor in the stream drains; see ptx.cuh).
template <typename... KArgs, typename... Args>
cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[3];
  int n = 0;
  if (g_num_windows.load(std::memory_order_relaxed) > 0) {
    L2Window w{nullptr, 0, 1.0f};
    {
      std::lock_guard<std::mutex> lk(g_win_mu);
      auto it = g_windows.find(st);
      if (it != g_windows.end()) w = it->second;
    }
    if (w.ptr != nullptr && w.bytes > 0) {
      attr[n].id = cudaLaunchAttributeAccessPolicyWindow;
      attr[n].val.accessPolicyWindow.base_ptr = w.ptr;
      attr[n].val.accessPolicyWindow.num_bytes = w.bytes;
      attr[n].val.accessPolicyWindow.hitRatio = w.ratio;
      attr[n].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr[n].val.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
      ++n;
    }
  }
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

inline int check_launch() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return static_cast<int>(cudaGetLastError());
}

// Tile width: minimise (rounds of the persistent schedule) x (tile width); ties go to the
// wider tile (fewer re-reads of A).  LDIT_GEMM_BN overrides for experiments.
std::atomic<int> g_forced_bn{-1};

std::atomic<int> g_cta_pair{-1};  // -1: read LDIT_GEMM_CTAS once; 1 or 2 afterwards

int gemm_ctas() {
  int v = g_cta_pair.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("LDIT_GEMM_CTAS");
    v = (e && atoi(e) == 1) ? 1 : 2;
    g_cta_pair.store(v);
  }
  return v;
}

// Tile width: minimise (rounds of the persistent schedule) x (tile width); ties go to the
// wider tile (fewer operand bytes per FLOP).  Widths that divide N are preferred.
// (Tried and rejected, same-box A/B: computing the ragged last row block -- 64 rows at M = 64 x 197 -- with "swapped"
// tiles C^T = W . A^T whose MMA N dimension carries the 64 tokens, so that QKV needs 8 schedule rounds instead of 9 and
// the residual GEMMs 2 x 256 instead of 3 x 192 columns.  A swapped tile streams 256 weight rows for a quarter of the
// math: at K = 768 it is L2-bandwidth-bound and takes ~0.7 of a regular tile, not 0.25-0.33, so the clusters that take
// the tails finish last: 2.72 vs 2.62 ms/step.)
int pick_bn(int M, int N, int ctas) {
  int forced = g_forced_bn.load(std::memory_order_relaxed);
  if (forced < 0) {
    const char* e = getenv("LDIT_GEMM_BN");
    forced = e ? atoi(e) : 0;
    g_forced_bn.store(forced);
  }
  if (forced == 128 || forced == 192 || forced == 256) return forced;
  const int units = num_sms() / ctas;  // CTAs or CTA pairs working in parallel
  const int mb = (M + kBM * ctas - 1) / (kBM * ctas);
  int best = 256;
  long best_cost = -1;
  const int cands[3] = {256, 192, 128};
  bool any_divides = false;
  for (int bn : cands) any_divides |= (N % bn == 0);
  for (int bn : cands) {
    if (any_divides && (N % bn)) continue;  // no ragged last column of tiles if it can be avoided
    const long tiles = static_cast<long>(mb) * ((N + bn - 1) / bn);
    const long rounds = (tiles + units - 1) / units;
    const long cost = rounds * (bn + 16);  // +16: per-tile fixed overhead in "column" units
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

template <int BN, int EPI, int CTAS>
int launch_gemm_maps(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, GemmArgs g, cudaStream_t st);

template <int BN, int EPI, int CTAS>
int launch_gemm_t(const void* A, const void* W, GemmArgs g, cudaStream_t st) {
  using Cfg = GemmCfg<BN, EPI, CTAS>;
  CUtensorMap tmA, tmB, tmC;
  int rc = make_tmap_bf16_2d(&tmA, A, g.M, g.K, kBM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, W, g.N, g.K, Cfg::B_ROWS);
  if (rc) return rc;
  if (EPI != EPI_PATCH) {  // epilogue staging: 32-row x 16-column boxes, rows of 32 B (bf16) / 64 B (fp32)
    rc = Cfg::OUT_F32 ? make_tmap_2d(&tmC, g.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g.M, g.N, g.ldo, 32, kEpiCols, CU_TENSOR_MAP_SWIZZLE_64B)
                      : make_tmap_2d(&tmC, g.out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.M, g.N, g.ldo, 32, kEpiCols, CU_TENSOR_MAP_SWIZZLE_32B);
    if (rc) return rc;
  } else {
    tmC = tmA;
  }
  g.num_m_blocks = (g.M + Cfg::TILE_M - 1) / Cfg::TILE_M;
  return launch_gemm_maps<BN, EPI, CTAS>(tmA, tmB, tmC, g, st);
}

template <int BN, int EPI, int CTAS>
int launch_gemm_maps(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, GemmArgs g, cudaStream_t st) {
  using Cfg = GemmCfg<BN, EPI, CTAS>;
  {
    cudaError_t e = ensure_smem(gemm_tcgen05_kernel<BN, EPI, CTAS>, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
#ifdef LDIT_DEBUG_HOOKS
  static const int dbg = [] { const char* e = getenv("LDIT_GEMM_DBG"); return e ? atoi(e) : 0; }();
  g.dbg = dbg;
  g.tl = g_gemm_tl;
#endif
  g.num_n_blocks = (g.N + BN - 1) / BN;
  const int tiles = g.num_m_blocks * g.num_n_blocks;
  const int units = num_sms() / CTAS;
  const int grid = (tiles < units ? tiles : units) * CTAS;
  cudaError_t e = launch_kernel(gemm_tcgen05_kernel<BN, EPI, CTAS>, dim3(grid), dim3(kGemmThreads), Cfg::SMEM_BYTES, st, CTAS, tmA, tmB, tmC, g);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return static_cast<int>(e); }
  return static_cast<int>(cudaGetLastError());
}

template <int EPI>
int launch_gemm(const void* A, const void* W, GemmArgs g, cudaStream_t st) {
  if (!A || !W || !g.out) return LDIT_E_NULL;
  if (g.M <= 0 || g.N <= 0 || g.K <= 0 || (g.N % 32) || (g.K % 8)) return LDIT_E_SHAPE;
  if (!aligned16(A) || !aligned16(W) || !aligned16(g.out) || !aligned16(g.bias) || !aligned16(g.scale) ||
      !aligned16(g.posb))
    return LDIT_E_ALIGN;
  const int ctas = gemm_ctas();
  const int bn = pick_bn(g.M, g.N, ctas);
  if (ctas == 2) {
    switch (bn) {
      case 128: return launch_gemm_t<128, EPI, 2>(A, W, g, st);
      case 192: return launch_gemm_t<192, EPI, 2>(A, W, g, st);
      default: return launch_gemm_t<256, EPI, 2>(A, W, g, st);
    }
  }
  switch (bn) {
    case 128: return launch_gemm_t<128, EPI, 1>(A, W, g, st);
    case 192: return launch_gemm_t<192, EPI, 1>(A, W, g, st);
    default: return launch_gemm_t<256, EPI, 1>(A, W, g, st);
  }
}

// 3x3 convolution as an implicit GEMM (EPI_CONV_BIAS / EPI_CONV_BIAS_F32): tensor maps over the channels-last images, patch shape
template <int BN, int EPI>
int launch_conv_t(const void* in, const void* w, GemmArgs g, int B, int H, int W, int Cin, cudaStream_t st) {
  using Cfg = GemmCfg<BN, EPI, 2>;
  CUtensorMap tmA, tmB, tmC;
  int rc = make_tmap_nhwc_4d(&tmA, in, B, H, W, Cin, 64, g.cv_tw, g.cv_th, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, w, g.N, g.K, Cfg::B_ROWS);
  if (rc) return rc;
  if (Cfg::OUT_F32) {   // fp32 output map: boxes of 16 channels = 64-byte rows, 64-byte swizzle (as the residual epilogue stages them)
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(g.N), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(B)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(g.N) * 4, static_cast<cuuint64_t>(W) * g.N * 4, static_cast<cuuint64_t>(H) * W * g.N * 4};
    cuuint32_t box[4] = {kEpiCols, static_cast<cuuint32_t>(g.cv_tw), static_cast<cuuint32_t>(32 / g.cv_tw), 1};
    rc = encode_cached(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g.out, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
  } else {
    rc = make_tmap_nhwc_4d(&tmC, g.out, B, H, W, g.N, kEpiCols, g.cv_tw, 32 / g.cv_tw, CU_TENSOR_MAP_SWIZZLE_32B);
  }
  if (rc) return rc;
  return launch_gemm_maps<BN, EPI, 2>(tmA, tmB, tmC, g, st);
}

template <typename T>
cudaError_t launch_pages_gather(const void* const* pages, const int* page_hw, __nv_bfloat16* a, int B, int H, int W, int max_page_w,
                                const float* ms, cudaStream_t st) {
  const int Gh = H / 16, Gw = W / 16;
  const T* const* pp = reinterpret_cast<const T* const*>(pages);
  // staging pitch: a multiple of 16 bytes; 6 rows must fit the 227 KB a block may opt into
  const int pitch = (max_page_w + 15) / 16 * 16;
  const size_t smem = static_cast<size_t>(6) * pitch * sizeof(T);
  if (max_page_w > 0 && smem <= 200 * 1024 && B <= 65535) {
    cudaError_t e = ensure_smem(pages_rows_im2col_kernel<T>, smem);
    if (e != cudaSuccess) return e;
    return launch_kernel(pages_rows_im2col_kernel<T>, dim3(H, B), dim3(128), smem, st, 1, pp, page_hw, a, H, W, Gh, Gw, ms[0], ms[1],
                         ms[2], ms[3], ms[4], ms[5], pitch);
  }
  const size_t threads = static_cast<size_t>(B) * 3 * H * (W / 8);
  const unsigned blocks = static_cast<unsigned>((threads + 255) / 256);
  return launch_kernel(pages_im2col_kernel<T>, dim3(blocks), dim3(256), 0, st, 1, pp, page_hw, a, B, H, W, Gh, Gw, ms[0], ms[1], ms[2],
                       ms[3], ms[4], ms[5]);
}

#ifdef LDIT_EXPERIMENTAL
// ---- fused MLP (mlp_fused.cuh): tile width and per-pair tile lists
int mlp_bn(int D, int I) {
  if (D % 192 == 0 && I % 192 == 0) return 192;
  if (D % 256 == 0 && I % 256 == 0) return 256;
  return 0;
}

// Every CTA pair gets its fc1 tiles first (dealt in row-block order, round by round), then its fc2 tiles.  The number
// of fc2 tiles per pair differs by at most one; a pair with one more fc2 tile gets w fewer fc1 tiles (w = cost of an
// fc2 tile in fc1 tiles), so that all lists take equally long; fc2 tiles go to the (pair, k-th fc2 slot) positions in
// order of their start time, i.e. the earliest free pairs take the earliest (long finished) row blocks.
int build_mlp_schedule(int M, int D, int I, int bn, int clusters, std::vector<int>& out) {
  const int mbs = (M + 2 * kBM - 1) / (2 * kBM), nb1 = I / bn, nb2 = D / bn;
  const int T1 = mbs * nb1, T2 = mbs * nb2, C = clusters;
  static const int mlp_dbg = [] { const char* e = getenv("LDIT_MLP_DBG"); return e ? atoi(e) : 0; }();
  if (mlp_dbg & 2) {   // experiments only: plain round-robin over fc1 then fc2, the order two separate launches produce
    std::vector<std::vector<int>> lists(C);
    for (int t = 0; t < T1; ++t) lists[t % C].push_back(t);
    for (int t = 0; t < T2; ++t) lists[t % C].push_back(T1 + t);
    size_t stride = 1;
    for (auto& l : lists) stride = std::max(stride, l.size() + 1);
    out.assign(static_cast<size_t>(C) * stride, -1);
    for (int c = 0; c < C; ++c) std::copy(lists[c].begin(), lists[c].end(), out.begin() + c * stride);
    return static_cast<int>(stride);
  }
  const double w = 1.15 * static_cast<double>(I) / D;
  std::vector<int> n2(C, T2 / C), q1(C, 0);
  for (int c = 0; c < T2 % C; ++c) n2[c] += 1;
  const double lstar = (T1 + w * T2) / C;
  long sum = 0;
  for (int c = 0; c < C; ++c) { q1[c] = std::max(0, static_cast<int>(lstar - w * n2[c] + 0.5)); sum += q1[c]; }
  while (sum > T1) {   // take from the longest list
    int best = -1;
    for (int c = 0; c < C; ++c) if (q1[c] > 0 && (best < 0 || q1[c] + w * n2[c] > q1[best] + w * n2[best])) best = c;
    if (best < 0) break;
    --q1[best]; --sum;
  }
  while (sum < T1) {   // give to the shortest list
    int best = 0;
    for (int c = 1; c < C; ++c) if (q1[c] + w * n2[c] < q1[best] + w * n2[best]) best = c;
    ++q1[best]; ++sum;
  }
  std::vector<std::vector<int>> lists(C);
  std::vector<int> given(C, 0);
  for (int next = 0; next < T1;) {
    for (int c = 0; c < C && next < T1; ++c)
      if (given[c] < q1[c]) { lists[c].push_back(next++); ++given[c]; }
  }
  struct Slot { double start; int c; };
  std::vector<Slot> slots;
  for (int c = 0; c < C; ++c)
    for (int k = 0; k < n2[c]; ++k) slots.push_back({q1[c] + k * w, c});
  std::stable_sort(slots.begin(), slots.end(), [](const Slot& a, const Slot& b) { return a.start < b.start; });
  for (int j = 0; j < T2; ++j) lists[slots[j].c].push_back(T1 + j);
  size_t stride = 1;
  for (auto& l : lists) stride = std::max(stride, l.size() + 1);
  out.assign(static_cast<size_t>(C) * stride, -1);
  for (int c = 0; c < C; ++c) std::copy(lists[c].begin(), lists[c].end(), out.begin() + c * stride);
  return static_cast<int>(stride);
}

template <int BN>
int launch_mlp_t(const void* a, const void* W1, void* h, const void* W2, void* x, MlpArgs g, cudaStream_t st) {
  using Cfg = MlpCfg<BN>;
  {
    cudaError_t e = ensure_smem(mlp_tcgen05_kernel<BN>, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  CUtensorMap tmA1, tmB1, tmC1, tmA2, tmB2, tmC2;
  int rc = make_tmap_bf16_2d(&tmA1, a, g.M, g.D, kBM);
  if (!rc) rc = make_tmap_bf16_2d(&tmB1, W1, g.I, g.D, Cfg::B_ROWS);
  if (!rc) rc = make_tmap_2d(&tmC1, h, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.M, g.I, g.I, 32, kEpiCols, CU_TENSOR_MAP_SWIZZLE_32B);
  if (!rc) rc = make_tmap_bf16_2d(&tmA2, h, g.M, g.I, kBM);
  if (!rc) rc = make_tmap_bf16_2d(&tmB2, W2, g.D, g.I, Cfg::B_ROWS);
  if (!rc) rc = make_tmap_2d(&tmC2, x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g.M, g.D, g.D, 32, kEpiCols, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  const int grid = (num_sms() / 2) * 2;
  cudaError_t e = launch_kernel(mlp_tcgen05_kernel<BN>, dim3(grid), dim3(kGemmThreads), Cfg::SMEM_BYTES, st, 2, tmA1, tmB1, tmC1, tmA2,
                                tmB2, tmC2, g);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return static_cast<int>(e); }
  return static_cast<int>(cudaGetLastError());
}

#endif  // LDIT_EXPERIMENTAL

// Patch embedding of 16-bit pixels with the A operand gathered by TMA (EPI_PATCH_TMA): ONE launch, no im2col scratch.
template <int BN>
int launch_patch_tma(const void* pixels, bool f16, const void* w, GemmArgs g, int B, int H, int W, cudaStream_t st) {
  using Cfg = GemmCfg<BN, EPI_PATCH_TMA, 2>;
  CUtensorMap tmA, tmB;
  // pixels [B, 3, H, W] as (px 16 | patch column | py 16 | patch row | image x channel), strides (bytes, dims 1..4) increasing;
  // one box = one pixel row of cv_tw x cv_th patches = 128 operand rows of 32 B
  cuuint64_t dims[5] = {16, static_cast<cuuint64_t>(g.pe_gw), 16, static_cast<cuuint64_t>(g.pe_gh), static_cast<cuuint64_t>(B) * 3};
  cuuint64_t strides[4] = {32, static_cast<cuuint64_t>(W) * 2, static_cast<cuuint64_t>(W) * 32, static_cast<cuuint64_t>(H) * W * 2};
  cuuint32_t box[5] = {16, static_cast<cuuint32_t>(g.cv_tw), 1, static_cast<cuuint32_t>(g.cv_th), 1};
  int rc = encode_cached(&tmA, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, pixels, dims, strides, box,
                         CU_TENSOR_MAP_SWIZZLE_32B);
  if (rc) return rc;
  rc = make_tmap_2d(&tmB, w, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.N, g.K, g.K, Cfg::B_ROWS, 64,
                    CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  return launch_gemm_maps<BN, EPI_PATCH_TMA, 2>(tmA, tmB, tmA, g, st);
}

// One launch with a TMA-gathered operand, or gather pass + CLS kernel + plain GEMM?  The TMA path tiles every image on
// its own (a pair tile is a 16 x 16 rectangle of the patch grid: 196 of 256 rows are real at 224 x 224), the plain GEMM runs
// over the dense [B P, 768] matrix but pays ~11 us for the gather pass.  Compare busiest-pair times in "columns"
// (rounds x tile width, as pick_bn) with the gather pass worth ~224 of them (calibrated: base224 29 vs 33 us,
// base512 57 vs 77 us for the TMA path, large224 45 vs 41 us against it).
bool patch_tma_preferred(int B, int H, int W, int D) {
  static const int force = [] { const char* e = getenv("LDIT_PATCH_TMA"); return e ? atoi(e) : -1; }();
  if (force == 0) return false;
  if (force == 2) return true;
  const int Gh = H / 16, Gw = W / 16, units = num_sms() / 2;
  auto cost = [&](long mblocks) {
    const int bn = pick_bn(static_cast<int>(mblocks * 256), D, 2);
    const long tiles = mblocks * ((D + bn - 1) / bn);
    return ((tiles + units - 1) / units) * static_cast<long>(bn);
  };
  long best = -1;
  const int tws[3] = {16, 8, 32};
  for (int tw : tws) {
    const int th = 128 / tw;
    const long tiles = static_cast<long>((Gw + tw - 1) / tw) * ((Gh + 2 * th - 1) / (2 * th));
    if (best < 0 || tiles < best) best = tiles;
  }
  const long dense = (static_cast<long>(B) * Gh * Gw + 255) / 256;
  return cost(static_cast<long>(B) * best) <= cost(dense) + 224;
}

int patch_embed_tma(const void* pixels, bool f16, const void* w, const void* pos_bias, const void* cls_pos, void* x, int B, int H, int W,
                    int D, cudaStream_t st) {
  const int Gh = H / 16, Gw = W / 16;
  GemmArgs g{};
  g.N = D; g.K = 768;
  g.out = x; g.ldo = D;
  g.P = Gh * Gw;
  g.posb = static_cast<const float*>(pos_bias);
  g.cls = static_cast<const float*>(cls_pos);
  g.pe_gw = Gw; g.pe_gh = Gh;
  g.a_f16 = f16 ? 1 : 0;
  // 128 patches per CTA as th rows x tw columns of the patch grid, the pair's second CTA below the first; fewest pair tiles wins
  long best = -1;
  const int tws[3] = {16, 8, 32};
  for (int tw : tws) {
    const int th = 128 / tw;
    const long tiles = static_cast<long>((Gw + tw - 1) / tw) * ((Gh + 2 * th - 1) / (2 * th));
    if (best < 0 || tiles < best) { best = tiles; g.cv_tw = tw; g.cv_th = th; }
  }
  g.cv_tx = (Gw + g.cv_tw - 1) / g.cv_tw;
  g.cv_ty = (Gh + 2 * g.cv_th - 1) / (2 * g.cv_th);
  g.num_m_blocks = B * g.cv_tx * g.cv_ty;
  g.M = g.num_m_blocks * 256;   // rows of the tiled index space (patches outside the grid included)
  const int bn = pick_bn(g.M, D, 2);
  switch (bn) {
    case 128: return launch_patch_tma<128>(pixels, f16, w, g, B, H, W, st);
    case 192: return launch_patch_tma<192>(pixels, f16, w, g, B, H, W, st);
    default: return launch_patch_tma<256>(pixels, f16, w, g, B, H, W, st);
  }
}

}  // namespace

// dW f32 [Nw, Kw] += dY^T A with dY bf16 [T, Nw] and A bf16 [T, Kw], both row-major as the forward / backward kernels
// wrote them: the GEMM's operands are MN-major in shared memory (no transposed copies), the token dimension T is the
// contraction and is cut into pieces that reduce-add into the same output tile (a [768, 768] gradient is 18 tiles for
// 74 CTA pairs otherwise).
template <int BN>
int launch_wgrad(const void* dY, const void* A, GemmArgs g, int T, cudaStream_t st) {
  using Cfg = GemmCfg<BN, EPI_WGRAD, 2>;
  CUtensorMap tmA, tmB, tmC;
  int rc = make_tmap_2d(&tmA, dY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, T, g.M, g.M, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tmap_2d(&tmB, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, T, g.N, g.N, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tmap_2d(&tmC, g.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g.M, g.N, g.ldo, 32, kEpiCols, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  g.num_m_blocks = (g.M + Cfg::TILE_M - 1) / Cfg::TILE_M;
  g.num_n_blocks = (g.N + BN - 1) / BN;
  const int nkb = (T + kBK - 1) / kBK;
  const int units = num_sms() / 2;
  const int mn = g.num_m_blocks * g.num_n_blocks;
  int want = (2 * units + mn - 1) / mn;                 // about two work items per CTA pair
  if (want < 1) want = 1;
  if (want > nkb / 8) want = nkb / 8 > 0 ? nkb / 8 : 1; // at least 8 k-blocks per piece: the pipeline fill is paid per piece
  g.kb_per_split = (nkb + want - 1) / want;
  g.ksplit = (nkb + g.kb_per_split - 1) / g.kb_per_split;   // every piece non-empty
  return launch_gemm_maps<BN, EPI_WGRAD, 2>(tmA, tmB, tmC, g, st);
}


// dA bf16 [M, Kin] = dY [M, Nout] x W [Nout, Kin], W read as nn.Linear stores it (MN-major B operand)
template <int BN>
int launch_dgrad(const void* dY, const void* W, GemmArgs g, cudaStream_t st) {
  using Cfg = GemmCfg<BN, EPI_DGRAD, 2>;
  CUtensorMap tmA, tmB, tmC;
  int rc = make_tmap_bf16_2d(&tmA, dY, g.M, g.K, kBM);
  if (rc) return rc;
  rc = make_tmap_2d(&tmB, W, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.K, g.N, g.N, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tmap_2d(&tmC, g.out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.M, g.N, g.ldo, 32, kEpiCols, CU_TENSOR_MAP_SWIZZLE_32B);
  if (rc) return rc;
  g.num_m_blocks = (g.M + Cfg::TILE_M - 1) / Cfg::TILE_M;
  return launch_gemm_maps<BN, EPI_DGRAD, 2>(tmA, tmB, tmC, g, st);
}

extern "C" {

int ldit_version(void) { return 100; }

const char* ldit_error_string(int code) {
  static thread_local char buf[128];
  switch (code) {
    case LDIT_OK: return "ok";
    case LDIT_E_NULL: return "required pointer is NULL";
    case LDIT_E_SHAPE: return "unsupported shape";
    case LDIT_E_ALIGN: return "pointer or pitch not 16-byte aligned";
    case LDIT_E_DTYPE: return "unknown dtype code";
    case LDIT_E_NO_DRIVER: return "cuTensorMapEncodeTiled unavailable (no CUDA driver)";
    case LDIT_E_UNSUPPORTED: return "entry point not compiled into this build (experimental kernel: rebuild with -DLDIT_EXPERIMENTAL)";
    default: break;
  }
  if (code >= 10000) {
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed with CUresult %d", code - 10000);
    return buf;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "unknown error";
}

void ldit_set_gemm_tile_n(int bn) { g_forced_bn.store((bn == 128 || bn == 192 || bn == 256) ? bn : 0); }
void ldit_debug_gemm_timeline(void* device_buffer) {
#ifdef LDIT_DEBUG_HOOKS
  g_gemm_tl = static_cast<long long*>(device_buffer);
#else
  (void)device_buffer;   // the product build carries no timeline hooks
#endif
}
void ldit_debug_attention_timeline(void* device_buffer) { g_attn_dbg = static_cast<long long*>(device_buffer); }
int ldit_has_experimental(void) {
#ifdef LDIT_EXPERIMENTAL
  return 1;
#else
  return 0;
#endif
}
void ldit_set_attention_impl(int impl) { g_attn_impl.store((impl == 1 || impl == 2 || impl == 4) ? impl : 0); }
void ldit_set_pdl(int on) { g_pdl.store(on ? 1 : 0); }

int ldit_set_l2_window(void* stream, void* ptr, size_t bytes, size_t set_aside_cap) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto drop = [&] {
    std::lock_guard<std::mutex> lk(g_win_mu);
    if (g_windows.erase(st)) g_num_windows.fetch_sub(1);
  };
  if (ptr == nullptr || bytes == 0) { drop(); return LDIT_OK; }
  int dev = 0, max_persist = 0, max_window = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
  if (e != cudaSuccess || max_persist <= 0 || max_window <= 0) {   // no persisting L2 on this device / partition: window stays off
    (void)cudaGetLastError();
    drop();
    return e != cudaSuccess ? static_cast<int>(e) : LDIT_OK;
  }
  size_t want = bytes < static_cast<size_t>(max_persist) ? bytes : static_cast<size_t>(max_persist);
  if (set_aside_cap > 0 && want > set_aside_cap) want = set_aside_cap;
  size_t limit = 0;
  {
    std::lock_guard<std::mutex> lk(g_dev_mu);
    DeviceState& d = dev_state_locked();
    if (want > d.persist_limit) {   // the set-aside is a device-wide limit: only ever grown, and only on request
      e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
      if (e != cudaSuccess) {   // e.g. under MPS: leave the window off, report, and let the caller carry on without it
        (void)cudaGetLastError();
        limit = static_cast<size_t>(-1);
      } else {
        d.persist_limit = want;
      }
    }
    if (limit == 0) limit = d.persist_limit;
  }
  if (limit == static_cast<size_t>(-1)) { drop(); return static_cast<int>(e); }
  L2Window w;
  w.ptr = ptr;
  w.bytes = bytes < static_cast<size_t>(max_window) ? bytes : static_cast<size_t>(max_window);
  // a window larger than the set-aside would thrash inside it: persist only the fraction that fits
  const size_t avail = (set_aside_cap > 0 && set_aside_cap < limit) ? set_aside_cap : limit;
  w.ratio = w.bytes <= avail ? 1.0f : static_cast<float>(avail) / static_cast<float>(w.bytes);
  std::lock_guard<std::mutex> lk(g_win_mu);
  if (g_windows.find(st) == g_windows.end()) g_num_windows.fetch_add(1);
  g_windows[st] = w;
  return LDIT_OK;
}

size_t ldit_workspace_bytes(int B, int H, int W, int D, int I) {
  if (B <= 0 || H < 16 || W < 16 || D <= 0 || I <= 0) return 0;
  const size_t M = static_cast<size_t>(B) * ((H / 16) * (W / 16) + 1);
  size_t wide = static_cast<size_t>(3) * D;
  if (static_cast<size_t>(I) > wide) wide = I;
  if (wide < 768) wide = 768;
  // residual stream x (f32 [M, D]) + LayerNorm output / context a (bf16 [M, D]) + wide buffer (bf16 [M, max(3D, I, 768)]:
  // QKV, MLP hidden, im2col scratch -- disjoint lifetimes); each rounded up to 1 KB so the three can share one allocation
  auto up = [](size_t v) { return (v + 1023) / 1024 * 1024; };
  return up(M * D * 4) + up(M * D * 2) + up(M * wide * 2);
}
void ldit_set_gemm_cta_pair(int ctas) { g_cta_pair.store(ctas == 1 ? 1 : 2); }

unsigned long long ldit_launch_count(void) { return g_launches.load(); }
void ldit_reset_launch_count(void) { g_launches.store(0); }

int ldit_layernorm(const void* x, const void* gamma, const void* beta, void* y, int rows, int D, float eps, void* stream) {
  if (!x || !gamma || !beta || !y) return LDIT_E_NULL;
  if (rows <= 0 || D <= 0 || (D % 128) || D > 2048) return LDIT_E_SHAPE;
  if (!aligned16(x) || !aligned16(gamma) || !aligned16(beta) || !aligned16(y)) return LDIT_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int lnw = ln_warps(D / 128);
  const int want = (rows + lnw - 1) / lnw;          // persistent: two blocks per SM, rows dealt evenly
  const int blocks = want < 2 * num_sms() ? want : 2 * num_sms();
  const float* xf = static_cast<const float*>(x);
  const float* gf = static_cast<const float*>(gamma);
  const float* bf = static_cast<const float*>(beta);
  __nv_bfloat16* yb = static_cast<__nv_bfloat16*>(y);
#define LDIT_LN_CASE(V) \
  case V: launch_kernel(layernorm_kernel<V>, dim3(blocks), dim3(lnw * 32), 0, st, 1, xf, gf, bf, yb, rows, eps); break;
  switch (D / 128) {
    LDIT_LN_CASE(1) LDIT_LN_CASE(2) LDIT_LN_CASE(3) LDIT_LN_CASE(4) LDIT_LN_CASE(5) LDIT_LN_CASE(6) LDIT_LN_CASE(7)
    LDIT_LN_CASE(8) LDIT_LN_CASE(9) LDIT_LN_CASE(10) LDIT_LN_CASE(11) LDIT_LN_CASE(12) LDIT_LN_CASE(13)
    LDIT_LN_CASE(14) LDIT_LN_CASE(15) LDIT_LN_CASE(16)
    default: return LDIT_E_SHAPE;
  }
#undef LDIT_LN_CASE
  return check_launch();
}

int ldit_add_layernorm(void* x, const void* branch, const void* gamma, const void* beta, void* y, int rows, int D, float eps,
                       void* stream) {
  if (!x || !branch || !gamma || !beta || !y) return LDIT_E_NULL;
  if (rows <= 0 || D <= 0 || (D % 128) || D > 2048) return LDIT_E_SHAPE;
  if (!aligned16(x) || !aligned16(branch) || !aligned16(gamma) || !aligned16(beta) || !aligned16(y)) return LDIT_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int lnw = aln_warps(D / 128);
  const int want = (rows + lnw - 1) / lnw;
  const int blocks = want < 2 * num_sms() ? want : 2 * num_sms();
  float* xf = static_cast<float*>(x);
  const __nv_bfloat16* bb = static_cast<const __nv_bfloat16*>(branch);
  const float* gf = static_cast<const float*>(gamma);
  const float* bf = static_cast<const float*>(beta);
  __nv_bfloat16* yb = static_cast<__nv_bfloat16*>(y);
#define LDIT_ALN_CASE(V) \
  case V: launch_kernel(add_layernorm_kernel<V>, dim3(blocks), dim3(lnw * 32), 0, st, 1, xf, bb, gf, bf, yb, rows, eps); break;
  switch (D / 128) {
    LDIT_ALN_CASE(1) LDIT_ALN_CASE(2) LDIT_ALN_CASE(3) LDIT_ALN_CASE(4) LDIT_ALN_CASE(5) LDIT_ALN_CASE(6) LDIT_ALN_CASE(7)
    LDIT_ALN_CASE(8) LDIT_ALN_CASE(9) LDIT_ALN_CASE(10) LDIT_ALN_CASE(11) LDIT_ALN_CASE(12) LDIT_ALN_CASE(13)
    LDIT_ALN_CASE(14) LDIT_ALN_CASE(15) LDIT_ALN_CASE(16)
    default: return LDIT_E_SHAPE;
  }
#undef LDIT_ALN_CASE
  return check_launch();
}

int ldit_gemm_bias_scale(const void* A, const void* W, const void* bias, const void* scale, void* out, int M, int N, int K,
                         void* stream) {
  GemmArgs g{};
  g.M = M; g.N = N; g.K = K;
  g.bias = static_cast<const float*>(bias);
  g.scale = static_cast<const float*>(scale);
  g.out = out; g.ldo = N;
  return launch_gemm<EPI_BIAS_SCALE>(A, W, g, static_cast<cudaStream_t>(stream));
}

int ldit_gemm_bias(const void* A, const void* W, const void* bias, void* out, int M, int N, int K, void* stream) {
  GemmArgs g{};
  g.M = M; g.N = N; g.K = K;
  g.bias = static_cast<const float*>(bias);
  g.out = out; g.ldo = N;
  return launch_gemm<EPI_BIAS>(A, W, g, static_cast<cudaStream_t>(stream));
}

int ldit_gemm_bias_gelu(const void* A, const void* W, const void* bias, void* out, int M, int N, int K, void* stream) {
  GemmArgs g{};
  g.M = M; g.N = N; g.K = K;
  g.bias = static_cast<const float*>(bias);
  g.out = out; g.ldo = N;
  return launch_gemm<EPI_BIAS_GELU>(A, W, g, static_cast<cudaStream_t>(stream));
}

int ldit_gemm_bias_scale_residual(const void* A, const void* W, const void* bias, const void* scale, void* x, int M, int N,
                                  int K, void* stream) {
  GemmArgs g{};
  g.M = M; g.N = N; g.K = K;
  g.bias = static_cast<const float*>(bias);
  g.scale = static_cast<const float*>(scale);
  g.out = x; g.ldo = N;
  // Experiment knob LDIT_FC2_REVERSE=1: a long-K residual GEMM (fc2) visits its row blocks last-to-first, so that of an
  // A operand the previous GEMM has just written front to back, and that is larger than what L2 keeps of it (77 MB at
  // base224), the rows still resident are consumed first.  Measured neutral (2.62 vs 2.64 ms/step): off.
  static const int rev = [] { const char* e = getenv("LDIT_FC2_REVERSE"); return e ? atoi(e) : 0; }();
  g.m_reverse = (rev && K > N) ? 1 : 0;
  g.round_bf16 = 1;   // same contribution, bit for bit, as ldit_gemm_bias_scale + ldit_add_layernorm
  return launch_gemm<EPI_SCALE_RESID>(A, W, g, static_cast<cudaStream_t>(stream));
}

int ldit_gemm_accumulate(const void* A, const void* W, void* acc, int M, int N, int K, void* stream) {
  GemmArgs g{};
  g.M = M; g.N = N; g.K = K;
  g.out = acc; g.ldo = N;
  return launch_gemm<EPI_SCALE_RESID>(A, W, g, static_cast<cudaStream_t>(stream));   // unrounded fp32 reduce-add
}

int ldit_gemm_dgrad(const void* dY, const void* W, void* dA, int M, int Nout, int Kin, void* stream) {
  if (!dY || !W || !dA) return LDIT_E_NULL;
  if (M <= 0 || Nout <= 0 || Kin <= 0 || (Nout % 8) || (Kin % 8)) return LDIT_E_SHAPE;
  if (!aligned16(dY) || !aligned16(W) || !aligned16(dA)) return LDIT_E_ALIGN;
  GemmArgs g{};
  g.M = M; g.N = Kin; g.K = Nout;
  g.out = dA; g.ldo = Kin;
  // tile widths whose per-CTA half is a whole number of 64-column atoms: 128 or 256, fewest schedule rounds x width
  const int units = num_sms() / 2, mb = (M + 255) / 256;
  const long c128 = ((static_cast<long>(mb) * ((Kin + 127) / 128) + units - 1) / units) * (128 + 16);
  const long c256 = ((static_cast<long>(mb) * ((Kin + 255) / 256) + units - 1) / units) * (256 + 16);
  if (c256 <= c128) return launch_dgrad<256>(dY, W, g, static_cast<cudaStream_t>(stream));
  return launch_dgrad<128>(dY, W, g, static_cast<cudaStream_t>(stream));
}

int ldit_gemm_wgrad(const void* dY, const void* A, void* dW, int T, int Nw, int Kw, void* stream) {
  if (!dY || !A || !dW) return LDIT_E_NULL;
  if (T <= 0 || Nw <= 0 || Kw <= 0 || (Nw % 8) || (Kw % 8)) return LDIT_E_SHAPE;
  if (!aligned16(dY) || !aligned16(A) || !aligned16(dW)) return LDIT_E_ALIGN;
  GemmArgs g{};
  g.M = Nw; g.N = Kw; g.K = T;
  g.out = dW; g.ldo = Kw;
  // 128-wide tiles unless the output is large enough to fill the machine with 256-wide ones
  const long tiles256 = static_cast<long>((Nw + 255) / 256) * ((Kw + 255) / 256);
  if (tiles256 >= num_sms() / 2) return launch_wgrad<256>(dY, A, g, T, static_cast<cudaStream_t>(stream));
  return launch_wgrad<128>(dY, A, g, T, static_cast<cudaStream_t>(stream));
}

#ifdef LDIT_EXPERIMENTAL
int ldit_mlp_clusters(void) { return num_sms() / 2; }

int ldit_mlp_schedule(int M, int D, int I, int* host_sched, int capacity) {
  if (M <= 0 || D <= 0 || I <= 0) return LDIT_E_SHAPE;
  const int bn = mlp_bn(D, I);
  if (!bn) return LDIT_E_SHAPE;
  std::vector<int> sched;
  const int stride = build_mlp_schedule(M, D, I, bn, ldit_mlp_clusters(), sched);
  if (host_sched != nullptr) {
    if (capacity < static_cast<int>(sched.size())) return LDIT_E_SHAPE;
    std::copy(sched.begin(), sched.end(), host_sched);
  }
  return stride;
}

int ldit_mlp_fused(const void* a, const void* W1, const void* b1, void* h, const void* W2, const void* b2, const void* lam2,
                   void* x, int M, int D, int I, const int* sched, int sched_stride, int* ready, void* stream) {
  if (!a || !W1 || !h || !W2 || !x || !sched || !ready) return LDIT_E_NULL;
  if (M <= 0 || D <= 0 || I <= 0 || (D % 8) || (I % 8) || sched_stride <= 0) return LDIT_E_SHAPE;
  if (!aligned16(a) || !aligned16(W1) || !aligned16(b1) || !aligned16(h) || !aligned16(W2) || !aligned16(b2) || !aligned16(lam2) ||
      !aligned16(x))
    return LDIT_E_ALIGN;
  const int bn = mlp_bn(D, I);
  if (!bn) return LDIT_E_SHAPE;
  MlpArgs g{};
  g.M = M; g.D = D; g.I = I;
  g.b1 = static_cast<const float*>(b1);
  g.b2 = static_cast<const float*>(b2);
  g.lam2 = static_cast<const float*>(lam2);
  g.nb1 = I / bn; g.nb2 = D / bn;
  g.num_m_blocks = (M + 2 * kBM - 1) / (2 * kBM);
  g.tiles1 = g.num_m_blocks * g.nb1;
  g.sched = sched; g.sched_stride = sched_stride;
  g.ready = ready;
  g.ready_target = 4 * 2 * I;
  static const int mlp_dbg = [] { const char* e = getenv("LDIT_MLP_DBG"); return e ? atoi(e) : 0; }();
  g.dbg = mlp_dbg;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (bn == 192) return launch_mlp_t<192>(a, W1, h, W2, x, g, st);
  return launch_mlp_t<256>(a, W1, h, W2, x, g, st);
}
#else   // !LDIT_EXPERIMENTAL: the fused fc1 + fc2 kernel (measured 14-21 % slower than two launches) is not built
int ldit_mlp_clusters(void) { return num_sms() / 2; }
int ldit_mlp_schedule(int, int, int, int*, int) { return LDIT_E_UNSUPPORTED; }
int ldit_mlp_fused(const void*, const void*, const void*, void*, const void*, const void*, const void*, void*, int, int, int,
                   const int*, int, int*, void*) { return LDIT_E_UNSUPPORTED; }
#endif

size_t ldit_patch_embed_scratch_bytes(int B, int H, int W) {
  if (B <= 0 || H < 16 || W < 16) return 0;
  return static_cast<size_t>(B) * (H / 16) * (W / 16) * 768 * 2;
}

int ldit_patch_embed(const void* pixels, int pixel_dtype, const void* w, const void* pos_bias, const void* cls_pos,
                     void* scratch, void* x, int B, int H, int W, int D, void* stream) {
  if (!pixels || !w || !pos_bias || !cls_pos || !scratch || !x) return LDIT_E_NULL;
  if (B <= 0 || H < 16 || W < 16 || (H % 16) || (W % 16) || D <= 0 || (D % 32)) return LDIT_E_SHAPE;
  if (!aligned16(pixels) || !aligned16(scratch) || !aligned16(x) || !aligned16(cls_pos)) return LDIT_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Gh = H / 16, Gw = W / 16, P = Gh * Gw;
  // 16-bit pixels (what the reference feeds on CUDA, R:trainer.py:155,168): TMA-fed im2col GEMM, one launch, scratch unused.
  // fp32 pixels keep the cast / gather pass below (TMA cannot convert, and a tf32 MMA would halve the tensor rate).
  if (gemm_ctas() == 2 && pixel_dtype == LDIT_DTYPE_BF16 && patch_tma_preferred(B, H, W, D))   // fp16 pixels: ldit_patch_embed_tma + fp16 weights
    return patch_embed_tma(pixels, false, w, pos_bias, cls_pos, x, B, H, W, D, st);
  const size_t threads = static_cast<size_t>(B) * 3 * H * (W / 8);
  const unsigned blocks = static_cast<unsigned>((threads + 255) / 256);
  __nv_bfloat16* a = static_cast<__nv_bfloat16*>(scratch);
  switch (pixel_dtype) {
    case LDIT_DTYPE_F32: launch_kernel(im2col_kernel<float>, dim3(blocks), dim3(256), 0, st, 1, static_cast<const float*>(pixels), a, B, H, W, Gh, Gw); break;
    case LDIT_DTYPE_F16: launch_kernel(im2col_kernel<__half>, dim3(blocks), dim3(256), 0, st, 1, static_cast<const __half*>(pixels), a, B, H, W, Gh, Gw); break;
    case LDIT_DTYPE_BF16:
      launch_kernel(im2col_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, 1, static_cast<const __nv_bfloat16*>(pixels), a, B, H, W, Gh, Gw);
      break;
    default: return LDIT_E_DTYPE;
  }
  int rc = check_launch();
  if (rc) return rc;
  launch_kernel(cls_rows_kernel, dim3((B * (D / 4) + 255) / 256), dim3(256), 0, st, 1, static_cast<const float*>(cls_pos), static_cast<float*>(x), B, P + 1, D);
  rc = check_launch();
  if (rc) return rc;
  GemmArgs g{};
  g.M = B * P; g.N = D; g.K = 768;
  g.out = x; g.ldo = D;
  g.P = P;
  g.posb = static_cast<const float*>(pos_bias);
  return launch_gemm<EPI_PATCH>(scratch, w, g, st);
}

int ldit_patch_embed_tma_preferred(int B, int H, int W, int D) {
  if (B <= 0 || H < 16 || W < 16 || D <= 0 || gemm_ctas() != 2) return 0;
  return patch_tma_preferred(B, H, W, D) ? 1 : 0;
}

int ldit_patch_embed_tma(const void* pixels, int pixel_dtype, const void* w, const void* pos_bias, const void* cls_pos, void* x,
                         int B, int H, int W, int D, void* stream) {
  if (!pixels || !w || !pos_bias || !cls_pos || !x) return LDIT_E_NULL;
  if (B <= 0 || H < 16 || W < 16 || (H % 16) || (W % 16) || D <= 0 || (D % 32)) return LDIT_E_SHAPE;
  if (pixel_dtype != LDIT_DTYPE_F16 && pixel_dtype != LDIT_DTYPE_BF16) return LDIT_E_DTYPE;
  if (!aligned16(pixels) || !aligned16(w) || !aligned16(x) || !aligned16(cls_pos) || !aligned16(pos_bias)) return LDIT_E_ALIGN;
  if (gemm_ctas() != 2) return LDIT_E_UNSUPPORTED;
  return patch_embed_tma(pixels, pixel_dtype == LDIT_DTYPE_F16, w, pos_bias, cls_pos, x, B, H, W, D, static_cast<cudaStream_t>(stream));
}

int ldit_patch_embed_pages(const void* const* pages, const int* page_hw, int max_page_w, int pixel_dtype, float mean0, float mean1,
                           float mean2, float std0, float std1, float std2, const void* w, const void* pos_bias, const void* cls_pos,
                           void* scratch, void* x, int B, int H, int W, int D, void* stream) {
  if (!pages || !page_hw || !w || !pos_bias || !cls_pos || !scratch || !x) return LDIT_E_NULL;
  if (B <= 0 || H < 16 || W < 16 || (H % 16) || (W % 16) || D <= 0 || (D % 32) || max_page_w < 0) return LDIT_E_SHAPE;
  if (!(std0 != 0.f) || !(std1 != 0.f) || !(std2 != 0.f)) return LDIT_E_SHAPE;
  if (!aligned16(scratch) || !aligned16(x) || !aligned16(cls_pos)) return LDIT_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int P = (H / 16) * (W / 16);
  __nv_bfloat16* a = static_cast<__nv_bfloat16*>(scratch);
  const float ms[6] = {mean0, mean1, mean2, std0, std1, std2};
  cudaError_t le;
  switch (pixel_dtype) {
    case LDIT_DTYPE_F32: le = launch_pages_gather<float>(pages, page_hw, a, B, H, W, max_page_w, ms, st); break;
    case LDIT_DTYPE_F16: le = launch_pages_gather<__half>(pages, page_hw, a, B, H, W, max_page_w, ms, st); break;
    case LDIT_DTYPE_BF16: le = launch_pages_gather<__nv_bfloat16>(pages, page_hw, a, B, H, W, max_page_w, ms, st); break;
    default: return LDIT_E_DTYPE;
  }
  if (le != cudaSuccess) { (void)cudaGetLastError(); return static_cast<int>(le); }
  int rc = check_launch();
  if (rc) return rc;
  launch_kernel(cls_rows_kernel, dim3((B * (D / 4) + 255) / 256), dim3(256), 0, st, 1, static_cast<const float*>(cls_pos), static_cast<float*>(x), B, P + 1, D);
  rc = check_launch();
  if (rc) return rc;
  GemmArgs g{};
  g.M = B * P; g.N = D; g.K = 768;
  g.out = x; g.ldo = D;
  g.P = P;
  g.posb = static_cast<const float*>(pos_bias);
  return launch_gemm<EPI_PATCH>(scratch, w, g, st);
}

static int attention_forward(const void* qkv, void* ctx, const void* bias_table, float* lse, int B, int N, int heads, int Gh, int Gw, void* stream);
int ldit_attention(const void* qkv, void* ctx, const void* bias_table, int B, int N, int heads, int Gh, int Gw, void* stream) {
  return attention_forward(qkv, ctx, bias_table, nullptr, B, N, heads, Gh, Gw, stream);
}
int ldit_attention_lse(const void* qkv, void* ctx, const void* bias_table, void* lse, int B, int N, int heads, int Gh, int Gw, void* stream) {
  if (!lse) return LDIT_E_NULL;
  if (!aligned16(lse)) return LDIT_E_ALIGN;
  return attention_forward(qkv, ctx, bias_table, static_cast<float*>(lse), B, N, heads, Gh, Gw, stream);
}
static int attention_forward(const void* qkv, void* ctx, const void* bias_table, float* lse, int B, int N, int heads, int Gh, int Gw, void* stream) {
  if (!qkv || !ctx) return LDIT_E_NULL;
  if (B <= 0 || heads <= 0 || Gh <= 0 || Gw <= 0 || N != Gh * Gw + 1) return LDIT_E_SHAPE;
  if (!aligned16(qkv) || !aligned16(ctx) || !aligned16(bias_table)) return LDIT_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int D = heads * 64;
  const int T = (2 * Gh - 1) * (2 * Gw - 1) + 3;
  const float scale_log2e = 0.125f * 1.4426950408889634f;
  int impl = g_attn_impl.load(std::memory_order_relaxed);
  if (impl < 0) {
    const char* e = getenv("LDIT_ATTN_IMPL");
    impl = e ? atoi(e) : 0;
    if (impl != 1 && impl != 2 && impl != 4) impl = 0;
    g_attn_impl.store(impl);
  }
  if (impl == 0) {
    AttnV3Args a{};
    a.bias_table = static_cast<const float*>(bias_table);
    a.B = B; a.N = N; a.heads = heads; a.D = D; a.Gh = Gh; a.Gw = Gw; a.T = T;
    a.n_ktiles = (N + kA3KT - 1) / kA3KT;
    a.n_qpairs = (N + 255) / 256;
    a.num_items = B * heads * a.n_qpairs;
    a.scale_log2e = scale_log2e;
    a.lse = lse;
    a.dbg = g_attn_dbg;
    CUtensorMap tmQ, tmKV, tmO;
    int rc = make_tmap_qkv_3d(&tmQ, qkv, B, N, 3 * D, 128);
    if (rc) return rc;
    rc = make_tmap_qkv_3d(&tmKV, qkv, B, N, 3 * D, kA3KT);
    if (rc) return rc;
    rc = make_tmap_rows_3d(&tmO, ctx, B, N, D, 32);   // ctx as [B, N, D]: box 32 rows x 64 cols
    if (rc) return rc;
    size_t smem = 1024 + kA3SmemTiles + (kA3NumBars + 2) * 8;
    if (bias_table) smem += (2 * static_cast<size_t>(a3_ext_len(Gh, Gw, T)) + static_cast<size_t>(a.n_ktiles) * kA3KT) * 4;
    if (smem > 227 * 1024) return LDIT_E_SHAPE;
    const int per_sm = smem <= 113 * 1024 ? 2 : 1;   // the kernel is sized to be resident twice per SM
    const int slots = per_sm * num_sms();
    const int grid = a.num_items < slots ? a.num_items : slots;
    // two CTAs per SM need a ~200 KB carve-out: ask for the largest one instead of leaving it to the driver's heuristic
    cudaError_t e = bias_table ? ensure_smem(attention_v3_kernel<true>, smem, true) : ensure_smem(attention_v3_kernel<false>, smem, true);
    if (e != cudaSuccess) return static_cast<int>(e);
    if (bias_table) launch_kernel(attention_v3_kernel<true>, dim3(grid), dim3(kA3Threads), smem, st, 1, tmQ, tmKV, tmO, a);
    else launch_kernel(attention_v3_kernel<false>, dim3(grid), dim3(kA3Threads), smem, st, 1, tmQ, tmKV, tmO, a);
    return check_launch();
  }
#ifndef LDIT_EXPERIMENTAL
  return LDIT_E_UNSUPPORTED;   // impl 1 / 2 / 4: the superseded round-1 kernels, only in -DLDIT_EXPERIMENTAL builds
#else
  if (lse) return LDIT_E_UNSUPPORTED;   // only the product kernel writes row statistics
  if (impl == 4) {
    AttnP2Args a{};
    a.ctx = static_cast<__nv_bfloat16*>(ctx);
    a.bias_table = static_cast<const float*>(bias_table);
    a.B = B; a.N = N; a.heads = heads; a.D = D; a.Gh = Gh; a.Gw = Gw; a.T = T;
    a.n_kv_tiles = (N + kA2MaxKv - 1) / kA2MaxKv;
    a.kv_tile = (((N + a.n_kv_tiles - 1) / a.n_kv_tiles) + 15) / 16 * 16;
    a.n_qpairs = (N + 255) / 256;
    a.num_items = B * heads * a.n_qpairs;
    a.scale_log2e = scale_log2e;
    a.dbg = g_attn_dbg;
    CUtensorMap tmQ, tmKV, tmO;
    int rc = make_tmap_qkv_3d(&tmQ, qkv, B, N, 3 * D, 128);
    if (rc) return rc;
    rc = make_tmap_qkv_3d(&tmKV, qkv, B, N, 3 * D, a.kv_tile);
    if (rc) return rc;
    rc = make_tmap_rows_3d(&tmO, ctx, B, N, D, 32);   // ctx as [B, N, D]: box 32 rows x 64 cols
    if (rc) return rc;
    size_t smem = 1024 + kA2SmemTiles + (kA2NumBars + 2) * 8;
    if (bias_table) smem += (2 * static_cast<size_t>(T) + N) * 4;
    if (smem > 227 * 1024) return LDIT_E_SHAPE;
    const int grid = a.num_items < num_sms() ? a.num_items : num_sms();
    const int nch = a.kv_tile / 16;
    const int bi = bias_table ? 1 : 0;
    cudaError_t e = cudaSuccess;
#define LDIT_ATTN_CASE(NCH)                                                                                             \
  case NCH:                                                                                                             \
    e = bi ? ensure_smem(attention_pp_kernel<true, NCH>, smem) : ensure_smem(attention_pp_kernel<false, NCH>, smem);     \
    if (e != cudaSuccess) return static_cast<int>(e);                                                                   \
    if (bi) launch_kernel(attention_pp_kernel<true, NCH>, dim3(grid), dim3(kA2Threads), smem, st, 1, tmQ, tmKV, tmO, a);                               \
    else launch_kernel(attention_pp_kernel<false, NCH>, dim3(grid), dim3(kA2Threads), smem, st, 1, tmQ, tmKV, tmO, a);                                 \
    break;
    switch (nch) {
      LDIT_ATTN_CASE(1) LDIT_ATTN_CASE(2) LDIT_ATTN_CASE(3) LDIT_ATTN_CASE(4)
      LDIT_ATTN_CASE(5) LDIT_ATTN_CASE(6)
      default: return LDIT_E_SHAPE;
    }
#undef LDIT_ATTN_CASE
    return check_launch();
  }
  if (impl == 2) {
    AttnTcArgs a{};
    a.ctx = static_cast<__nv_bfloat16*>(ctx);
    a.bias_table = static_cast<const float*>(bias_table);
    a.B = B; a.N = N; a.heads = heads; a.D = D; a.Gh = Gh; a.Gw = Gw; a.T = T;
    a.n_kv_tiles = (N + kAtcMaxKv - 1) / kAtcMaxKv;
    a.kv_tile = (((N + a.n_kv_tiles - 1) / a.n_kv_tiles) + 15) / 16 * 16;
    a.scale_log2e = scale_log2e;
    CUtensorMap tmQ, tmKV;
    int rc = make_tmap_qkv_3d(&tmQ, qkv, B, N, 3 * D, kAtcQ);
    if (rc) return rc;
    rc = make_tmap_qkv_3d(&tmKV, qkv, B, N, 3 * D, a.kv_tile);
    if (rc) return rc;
    size_t smem = 1024 + 16384 + 2 * kAtcMaxKv * 128 + 64;
    if (bias_table) smem += (static_cast<size_t>(T) + N) * 4;
    if (smem > 110 * 1024) return LDIT_E_SHAPE;
    dim3 grid((N + kAtcQ - 1) / kAtcQ, heads, B);
    {
      cudaError_t e = bias_table ? ensure_smem(attention_tc_kernel<true>, smem) : ensure_smem(attention_tc_kernel<false>, smem);
      if (e != cudaSuccess) return static_cast<int>(e);
    }
    if (bias_table) launch_kernel(attention_tc_kernel<true>, dim3(grid), dim3(kAtcThreads), smem, st, 1, tmQ, tmKV, a);
    else launch_kernel(attention_tc_kernel<false>, dim3(grid), dim3(kAtcThreads), smem, st, 1, tmQ, tmKV, a);
    return check_launch();
  }
  AttnArgs a{};
  a.qkv = static_cast<const __nv_bfloat16*>(qkv);
  a.ctx = static_cast<__nv_bfloat16*>(ctx);
  a.bias_table = static_cast<const float*>(bias_table);
  a.B = B; a.N = N; a.heads = heads; a.D = D;
  a.Gh = Gh; a.Gw = Gw; a.T = T;
  a.scale_log2e = scale_log2e;
  dim3 grid((N + kAttBQ - 1) / kAttBQ, heads, B);
  size_t smem = 8192 + 32768;
  if (bias_table) {
    smem += static_cast<size_t>(a.T) * 4 + static_cast<size_t>(N) * 4;
    if (smem > 227 * 1024) return LDIT_E_SHAPE;
    {
      cudaError_t e = ensure_smem(attention_mma_kernel<true>, smem);
      if (e != cudaSuccess) return static_cast<int>(e);
    }
    launch_kernel(attention_mma_kernel<true>, dim3(grid), dim3(128), smem, st, 1, a);
  } else {
    launch_kernel(attention_mma_kernel<false>, dim3(grid), dim3(128), smem, st, 1, a);
  }
  return check_launch();
#endif
}

int ldit_resample_taps(const void* x, void* out, int B, int Gh, int Gw, int D, float scale, void* stream) {
  if (!x || !out) return LDIT_E_NULL;
  if (B <= 0 || Gh <= 0 || Gw <= 0 || D <= 0 || (D % 8) || !(scale > 0.f)) return LDIT_E_SHAPE;
  if (!aligned16(x) || !aligned16(out)) return LDIT_E_ALIGN;
  const int oh = static_cast<int>(floorf(Gh * scale)), ow = static_cast<int>(floorf(Gw * scale));
  if (oh <= 0 || ow <= 0) return LDIT_E_SHAPE;
  if (D / 8 * kTapPix > 1024 || B > 65535) return LDIT_E_SHAPE;
  if (scale == 1.0f) {  // no interpolation (R:57 skips it): CLS-less fp32 -> bf16 copy
    const size_t n = static_cast<size_t>(B) * Gh * Gw * (D / 8);
    launch_kernel(cast_tokens_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 1,
                  static_cast<const float*>(x), static_cast<__nv_bfloat16*>(out), B, Gh * Gw, D);
    return check_launch();
  }
  if ((scale == 4.0f || scale == 2.0f) && D / 8 * kUpCells <= 512) {  // cell-based integer up-sampling
    dim3 ublock(D / 8, kUpCells);
    dim3 ugrid((Gh * Gw + kUpCells - 1) / kUpCells, B);
    const float* xf = static_cast<const float*>(x);
    __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(out);
    if (scale == 4.0f) launch_kernel(upsample_taps_kernel<4>, ugrid, ublock, 0, static_cast<cudaStream_t>(stream), 1, xf, ob, Gh * Gw + 1, D, Gh, Gw);
    else launch_kernel(upsample_taps_kernel<2>, ugrid, ublock, 0, static_cast<cudaStream_t>(stream), 1, xf, ob, Gh * Gw + 1, D, Gh, Gw);
    return check_launch();
  }
  dim3 block(D / 8, kTapPix);
  dim3 grid((oh * ow + kTapPix - 1) / kTapPix, B);
  launch_kernel(resample_taps_kernel, grid, block, 0, static_cast<cudaStream_t>(stream), 1, static_cast<const float*>(x), static_cast<__nv_bfloat16*>(out), Gh * Gw + 1, D, Gh, Gw, oh, ow, 1.0f / scale);
  return check_launch();
}

int ldit_fpn_merge(const void* lat, const void* top, void* out, int B, int Gh, int Gw, int C, float scale, int top_h,
                   int top_w, void* stream) {
  if (!lat || !out) return LDIT_E_NULL;
  if (B <= 0 || B > 65535 || Gh <= 0 || Gw <= 0 || C <= 0 || (C % 8) || C / 8 > 1024 || !(scale > 0.f)) return LDIT_E_SHAPE;
  if (top && (top_h <= 0 || top_w <= 0)) return LDIT_E_SHAPE;
  if (!aligned16(lat) || !aligned16(top) || !aligned16(out)) return LDIT_E_ALIGN;
  const int oh = static_cast<int>(floorf(Gh * scale)), ow = static_cast<int>(floorf(Gw * scale));
  if (oh <= 0 || ow <= 0) return LDIT_E_SHAPE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* lb = static_cast<const __nv_bfloat16*>(lat);
  const __nv_bfloat16* tb = static_cast<const __nv_bfloat16*>(top);
  __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(out);
  if (top && (scale == 4.0f || scale == 2.0f) && top_h * 2 == oh && top_w * 2 == ow && C / 8 <= 256) {  // cell-based up-sampling
    const int cells = (256 / (C / 8)) < 4 ? (256 / (C / 8)) : 4;
    dim3 ublock(C / 8, cells);
    dim3 ugrid((Gh * Gw + cells - 1) / cells, B);
    if (scale == 4.0f) launch_kernel(fpn_merge_up_kernel<4>, ugrid, ublock, 0, st, 1, lb, tb, ob, C, Gh, Gw);
    else launch_kernel(fpn_merge_up_kernel<2>, ugrid, ublock, 0, st, 1, lb, tb, ob, C, Gh, Gw);
    return check_launch();
  }
  const int pix = (1024 / (C / 8)) < 8 ? (1024 / (C / 8)) : 8;   // output pixels per block
  dim3 block(C / 8, pix);
  dim3 grid((oh * ow + pix - 1) / pix, B);
  launch_kernel(fpn_merge_kernel, grid, block, 0, static_cast<cudaStream_t>(stream), 1, static_cast<const __nv_bfloat16*>(lat),
                static_cast<const __nv_bfloat16*>(top), static_cast<__nv_bfloat16*>(out), C, Gh, Gw, oh, ow, 1.0f / scale, top_h, top_w);
  return check_launch();
}

static int conv3x3_impl(const void* in, const void* w, const void* bias, void* out, int B, int H, int W, int Cin, int Cout,
                        bool out_f32, void* stream) {
  if (!in || !w || !out) return LDIT_E_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || (Cin % 64) || Cout <= 0 || (Cout % 128)) return LDIT_E_SHAPE;
  if (!aligned16(in) || !aligned16(w) || !aligned16(bias) || !aligned16(out)) return LDIT_E_ALIGN;
  GemmArgs g{};
  g.N = Cout; g.K = 9 * Cin;
  g.bias = static_cast<const float*>(bias);
  g.out = out; g.ldo = Cout;
  // patch shape: 128 output pixels per CTA as th x tw, a pair covers th x 2tw; fewest pair tiles wins
  long best = -1;
  const int tws[3] = {16, 8, 32};
  for (int tw : tws) {
    const int th = 128 / tw;
    const long tiles = static_cast<long>((W + 2 * tw - 1) / (2 * tw)) * ((H + th - 1) / th);
    if (best < 0 || tiles < best) { best = tiles; g.cv_tw = tw; g.cv_th = th; }
  }
  g.cv_tx = (W + 2 * g.cv_tw - 1) / (2 * g.cv_tw);
  g.cv_ty = (H + g.cv_th - 1) / g.cv_th;
  g.cv_cblocks = Cin / 64;
  g.num_m_blocks = B * g.cv_tx * g.cv_ty;
  g.M = g.num_m_blocks * 256;   // rows of the implicit GEMM including the pixels past the image edge
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int bn = (Cout % 256 == 0) ? pick_bn(g.M, Cout, 2) : 128;
  if (out_f32) {
    if (bn == 256) return launch_conv_t<256, EPI_CONV_BIAS_F32>(in, w, g, B, H, W, Cin, st);
    return launch_conv_t<128, EPI_CONV_BIAS_F32>(in, w, g, B, H, W, Cin, st);
  }
  if (bn == 256) return launch_conv_t<256, EPI_CONV_BIAS>(in, w, g, B, H, W, Cin, st);
  return launch_conv_t<128, EPI_CONV_BIAS>(in, w, g, B, H, W, Cin, st);
}

int ldit_conv3x3_bias(const void* in, const void* w, const void* bias, void* out, int B, int H, int W, int Cin, int Cout,
                      void* stream) {
  return conv3x3_impl(in, w, bias, out, B, H, W, Cin, Cout, false, stream);
}

int ldit_conv3x3_bias_f32(const void* in, const void* w, const void* bias, void* out, int B, int H, int W, int Cin, int Cout,
                          void* stream) {
  return conv3x3_impl(in, w, bias, out, B, H, W, Cin, Cout, true, stream);
}

// ------------------------------------------------------------------ backward of one BeitLayer (first slice)
// Plain stream-ordered launches (no programmatic dependent launch: these kernels do not carry griddepcontrol).
int ldit_transpose_bf16(const void* in, void* out, int R, int C, int ld_out, void* stream) {
  if (!in || !out) return LDIT_E_NULL;
  if (R <= 0 || C <= 0 || ld_out < R) return LDIT_E_SHAPE;
  transpose_bf16_kernel<<<dim3((C + 63) / 64, (R + 63) / 64), dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), R, C, ld_out);
  return check_launch();
}

int ldit_colsum_bf16(const void* in, void* out, int R, int C, int ld, void* stream) {
  if (!in || !out) return LDIT_E_NULL;
  if (R <= 0 || C <= 0 || (C % 2) || ld < C || (ld % 2)) return LDIT_E_SHAPE;
  const int rpb = 256;
  colsum_bf16_kernel<<<dim3((C + 63) / 64, (R + rpb - 1) / rpb), dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), static_cast<float*>(out), R, C, ld, rpb);
  return check_launch();
}

int ldit_gelu(const void* pre, void* h, size_t n, void* stream) {
  if (!pre || !h) return LDIT_E_NULL;
  if (n == 0 || (n % 8)) return LDIT_E_SHAPE;
  if (!aligned16(pre) || !aligned16(h)) return LDIT_E_ALIGN;
  gelu_fwd_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(pre), static_cast<__nv_bfloat16*>(h), n / 8);
  return check_launch();
}

int ldit_gelu_bwd(const void* dh, const void* pre, void* dpre, size_t n, void* stream) {
  if (!dh || !pre || !dpre) return LDIT_E_NULL;
  if (n == 0 || (n % 8)) return LDIT_E_SHAPE;
  if (!aligned16(dh) || !aligned16(pre) || !aligned16(dpre)) return LDIT_E_ALIGN;
  gelu_bwd_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dh), static_cast<const __nv_bfloat16*>(pre), static_cast<__nv_bfloat16*>(dpre), n / 8);
  return check_launch();
}

int ldit_scale_residual_rows(const void* x, const void* branch, const void* lam, const void* row_scale, int rows_per_image, void* y,
                             int rows, int D, void* stream) {
  if (!x || !branch || !y) return LDIT_E_NULL;
  if (rows <= 0 || D <= 0 || (D % 8) || (row_scale && rows_per_image <= 0)) return LDIT_E_SHAPE;
  if (!aligned16(x) || !aligned16(branch) || !aligned16(lam) || !aligned16(y)) return LDIT_E_ALIGN;
  const size_t n8 = static_cast<size_t>(rows) * (D / 8);
  scale_residual_fwd_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float*>(x), static_cast<const __nv_bfloat16*>(branch), static_cast<const float*>(lam),
      static_cast<const float*>(row_scale), rows_per_image, static_cast<float*>(y), rows, D);
  return check_launch();
}
int ldit_scale_residual(const void* x, const void* branch, const void* lam, void* y, int rows, int D, void* stream) {
  return ldit_scale_residual_rows(x, branch, lam, nullptr, 1, y, rows, D, stream);
}

int ldit_scale_residual_rows_bwd(const void* dy, const void* branch, const void* lam, const void* row_scale, int rows_per_image,
                                 void* dbranch, void* dlam, int rows, int D, void* stream) {
  if (!dy || !branch || !dbranch) return LDIT_E_NULL;
  if (rows <= 0 || D <= 0 || (D % 8) || (row_scale && rows_per_image <= 0)) return LDIT_E_SHAPE;
  if (!aligned16(dy) || !aligned16(branch) || !aligned16(lam) || !aligned16(dbranch) || !aligned16(dlam)) return LDIT_E_ALIGN;
  const int rpb = 64, threads = (D / 8 < 128) ? D / 8 : 128;
  scale_residual_bwd_kernel<<<dim3((D / 8 + threads - 1) / threads, (rows + rpb - 1) / rpb), dim3(threads, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float*>(dy), static_cast<const __nv_bfloat16*>(branch), static_cast<const float*>(lam),
      static_cast<const float*>(row_scale), rows_per_image, static_cast<__nv_bfloat16*>(dbranch), static_cast<float*>(dlam), rows, D, rpb);
  return check_launch();
}
int ldit_scale_residual_bwd(const void* dy, const void* branch, const void* lam, void* dbranch, void* dlam, int rows, int D,
                            void* stream) {
  return ldit_scale_residual_rows_bwd(dy, branch, lam, nullptr, 1, dbranch, dlam, rows, D, stream);
}

int ldit_layernorm_bwd(const void* x, const void* gamma, const void* dy, const void* dx_in, void* dx_out, void* dgamma, void* dbeta,
                       int rows, int D, float eps, void* stream) {
  if (!x || !gamma || !dy || !dx_out || !dgamma || !dbeta) return LDIT_E_NULL;
  if (rows <= 0 || D <= 0 || (D % 128) || D > 2048) return LDIT_E_SHAPE;
  if (!aligned16(x) || !aligned16(gamma) || !aligned16(dy) || !aligned16(dx_in) || !aligned16(dx_out)) return LDIT_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int want = (rows + 7) / 8;
  const int blocks = want < 4 * num_sms() ? want : 4 * num_sms();
#define LDIT_LNB_CASE(V)                                                                                                        \
  case V:                                                                                                                       \
    layernorm_bwd_kernel<V><<<blocks, 256, 0, st>>>(static_cast<const float*>(x), static_cast<const float*>(gamma),            \
                                                    static_cast<const __nv_bfloat16*>(dy), static_cast<const float*>(dx_in),    \
                                                    static_cast<float*>(dx_out), static_cast<float*>(dgamma),                  \
                                                    static_cast<float*>(dbeta), rows, eps);                                    \
    break;
  switch (D / 128) {
    LDIT_LNB_CASE(1) LDIT_LNB_CASE(2) LDIT_LNB_CASE(3) LDIT_LNB_CASE(4) LDIT_LNB_CASE(5) LDIT_LNB_CASE(6) LDIT_LNB_CASE(7)
    LDIT_LNB_CASE(8) LDIT_LNB_CASE(9) LDIT_LNB_CASE(10) LDIT_LNB_CASE(11) LDIT_LNB_CASE(12) LDIT_LNB_CASE(13)
    LDIT_LNB_CASE(14) LDIT_LNB_CASE(15) LDIT_LNB_CASE(16)
    default: return LDIT_E_SHAPE;
  }
#undef LDIT_LNB_CASE
  return check_launch();
}

int ldit_attention_bwd(const void* qkv, const void* dctx, void* dqkv, int B, int N, int heads, void* stream) {
  if (!qkv || !dctx || !dqkv) return LDIT_E_NULL;
  if (B <= 0 || heads <= 0 || N <= 0) return LDIT_E_SHAPE;
  if (N > 256) return LDIT_E_UNSUPPORTED;   // two 128-row query tiles x two 128-key halves per CTA (224 x 224 pages: 197 tokens)
  if (!aligned16(qkv) || !aligned16(dctx) || !aligned16(dqkv)) return LDIT_E_ALIGN;
  const int D = heads * 64;
  AttnBwdArgs a{};
  a.qkv = static_cast<const __nv_bfloat16*>(qkv);
  a.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  a.B = B; a.N = N; a.heads = heads; a.D = D;
  a.scale = 0.125f;
  a.scale_log2e = 0.125f * 1.4426950408889634f;
  CUtensorMap tmQKV, tmDO;
  int rc = make_tmap_qkv_3d(&tmQKV, qkv, B, N, 3 * D, 128);
  if (rc) return rc;
  rc = make_tmap_rows_3d(&tmDO, dctx, B, N, D, 128);
  if (rc) return rc;
  const size_t smem = 1024 + kAbtSmemTiles + 64;
  cudaError_t e = ensure_smem(attention_bwd_tc_kernel, smem, false);
  if (e != cudaSuccess) return static_cast<int>(e);
  attention_bwd_tc_kernel<<<B * heads, kAbtThreads, smem, static_cast<cudaStream_t>(stream)>>>(tmQKV, tmDO, a);
  return check_launch();
}

int ldit_attention_bwd_flash(const void* qkv, const void* ctx, const void* lse, const void* dctx, void* dqkv, void* dq_acc, void* delta,
                             const void* bias_table, void* dbias, int B, int N, int heads, int Gh, int Gw, void* stream) {
  if (!qkv || !ctx || !lse || !dctx || !dqkv || !dq_acc || !delta) return LDIT_E_NULL;
  if (B <= 0 || heads <= 0 || N <= 0 || B * heads > 65535) return LDIT_E_SHAPE;
  if ((bias_table != nullptr) != (dbias != nullptr)) return LDIT_E_NULL;
  if (bias_table && (Gh <= 0 || Gw <= 0 || N != Gh * Gw + 1)) return LDIT_E_SHAPE;
  if (!aligned16(bias_table) || !aligned16(dbias)) return LDIT_E_ALIGN;
  if (!aligned16(qkv) || !aligned16(ctx) || !aligned16(lse) || !aligned16(dctx) || !aligned16(dqkv) || !aligned16(dq_acc) || !aligned16(delta))
    return LDIT_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int D = heads * 64;
  const size_t rows = static_cast<size_t>(B) * N;
  cudaError_t e = cudaMemsetAsync(dq_acc, 0, rows * D * sizeof(float), st);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return static_cast<int>(e); }
  attention_delta_kernel<<<static_cast<unsigned>((rows * heads + 255) / 256), 256, 0, st>>>(
      static_cast<const __nv_bfloat16*>(dctx), static_cast<const __nv_bfloat16*>(ctx), static_cast<float*>(delta), B, N, heads);
  int rc = check_launch();
  if (rc) return rc;
  AttnBwdFlashArgs a{};
  a.lse = static_cast<const float*>(lse);
  a.delta = static_cast<const float*>(delta);
  a.dq_acc = static_cast<float*>(dq_acc);
  a.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  a.B = B; a.N = N; a.heads = heads; a.D = D;
  a.nqt = (N + 127) / 128;
  a.scale = 0.125f;
  a.scale_log2e = 0.125f * 1.4426950408889634f;
  CUtensorMap tmQKV, tmDO;
  rc = make_tmap_qkv_3d(&tmQKV, qkv, B, N, 3 * D, 128);
  if (rc) return rc;
  rc = make_tmap_rows_3d(&tmDO, dctx, B, N, D, 128);
  if (rc) return rc;
  size_t smem = 1024 + kAbfSmemTiles + 64;
  if (bias_table) {
    a.bias_table = static_cast<const float*>(bias_table);
    a.dbias = static_cast<float*>(dbias);
    a.Gh = Gh; a.Gw = Gw; a.T = (2 * Gh - 1) * (2 * Gw - 1) + 3;
    smem += (2 * static_cast<size_t>(a.T) + 128) * 4;
    if (smem > 227 * 1024) return LDIT_E_SHAPE;
    e = ensure_smem(attention_bwd_flash_kernel<true>, smem, false);
    if (e != cudaSuccess) return static_cast<int>(e);
    attention_bwd_flash_kernel<true><<<dim3(a.nqt, B * heads), kAbfThreads, smem, st>>>(tmQKV, tmDO, a);
  } else {
    e = ensure_smem(attention_bwd_flash_kernel<false>, smem, false);
    if (e != cudaSuccess) return static_cast<int>(e);
    attention_bwd_flash_kernel<false><<<dim3(a.nqt, B * heads), kAbfThreads, smem, st>>>(tmQKV, tmDO, a);
  }
  rc = check_launch();
  if (rc) return rc;
  dq_cast_kernel<<<static_cast<unsigned>((rows * (D / 8) + 255) / 256), 256, 0, st>>>(static_cast<const float*>(dq_acc), static_cast<__nv_bfloat16*>(dqkv), rows, D);
  return check_launch();
}

int ldit_resample_taps_bwd(const void* dout, void* dx, int B, int Gh, int Gw, int D, float scale, void* stream) {
  if (!dout || !dx) return LDIT_E_NULL;
  if (B <= 0 || Gh <= 0 || Gw <= 0 || D <= 0 || (D % 8) || !(scale > 0.f)) return LDIT_E_SHAPE;
  if (!aligned16(dout) || !aligned16(dx)) return LDIT_E_ALIGN;
  const int oh = static_cast<int>(floorf(Gh * scale)), ow = static_cast<int>(floorf(Gw * scale));
  if (oh <= 0 || ow <= 0 || D / 8 * kTapPix > 1024 || B > 65535) return LDIT_E_SHAPE;
  dim3 block(D / 8, kTapPix);
  dim3 grid((Gh * Gw + kTapPix - 1) / kTapPix, B);
  resample_taps_bwd_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dout), static_cast<float*>(dx), Gh * Gw + 1, D, Gh, Gw, oh, ow, 1.0f / scale, scale);
  return check_launch();
}

int ldit_batch_sum(const void* x, void* out, int B, int R, void* stream) {
  if (!x || !out) return LDIT_E_NULL;
  if (B <= 0 || R <= 0 || (R % 4)) return LDIT_E_SHAPE;
  if (!aligned16(x) || !aligned16(out)) return LDIT_E_ALIGN;
  batch_sum_kernel<<<(R / 4 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const float*>(x), static_cast<float*>(out), B, R);
  return check_launch();
}

int ldit_resize_rows(const void* src, void* dst, const void* add, int h, int w, int oh, int ow, int C, int bicubic, void* stream) {
  if (!src || !dst) return LDIT_E_NULL;
  if (h <= 0 || w <= 0 || oh <= 0 || ow <= 0 || C <= 0) return LDIT_E_SHAPE;
  const size_t n = static_cast<size_t>(oh) * ow * C;
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float* s = static_cast<const float*>(src);
  const float* a = static_cast<const float*>(add);
  float* d = static_cast<float*>(dst);
  // plain launches (no programmatic serialization): weight preparation, outside the forward's launch chain
  if (bicubic) resize_rows_kernel<true><<<blocks, 256, 0, st>>>(s, d, a, h, w, oh, ow, C);
  else resize_rows_kernel<false><<<blocks, 256, 0, st>>>(s, d, a, h, w, oh, ow, C);
  return check_launch();
}

static int subsample2_impl(const void* in, void* out, int B, int H, int W, int C, int esize, void* stream) {
  if (!in || !out) return LDIT_E_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || ((C * esize) % 16)) return LDIT_E_SHAPE;
  if (!aligned16(in) || !aligned16(out)) return LDIT_E_ALIGN;
  const int c16 = C * esize / 16;
  const size_t n = static_cast<size_t>(B) * ((H + 1) / 2) * ((W + 1) / 2) * c16;
  launch_kernel(subsample2_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 1,
                static_cast<const uint4*>(in), static_cast<uint4*>(out), B, H, W, c16);
  return check_launch();
}

int ldit_subsample2(const void* in, void* out, int B, int H, int W, int C, void* stream) {
  return subsample2_impl(in, out, B, H, W, C, 2, stream);
}

int ldit_subsample2_f32(const void* in, void* out, int B, int H, int W, int C, void* stream) {
  return subsample2_impl(in, out, B, H, W, C, 4, stream);
}

}  // extern "C"
