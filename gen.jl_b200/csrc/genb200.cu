// genb200.cu -- C ABI (include/genb200.h) over the sm_100a kernels: the device-resident
// structure-of-arrays ParticleFilterState (particle_filter.jl:35-41) and Gen's particle-filter /
// importance-sampling entry points for the supported model class.
//
// There is no CPU fallback anywhere in this file: every compute entry needs a CUDA device and
// fails with GB_ENOGPU otherwise.
#include "../../include/genb200.h"
#include "gbmath.cuh"
#include "kernels_resample.cuh"
#include "kernels_spec.cuh"
#include "kernels_mh.cuh"
#include "kernels_small.cuh"
#include "kernels_mid.cuh"
#include "kernels_msort.cuh"

#include <algorithm>
#include <atomic>
#include <mutex>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <type_traits>
#include <unordered_map>
#include <vector>

using namespace gb;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail(e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver ? GB_ENOGPU : GB_ECUDA, \
                  std::string(#call) + ": " + cudaGetErrorString(e__));                            \
  } while (0)
#define REQ(cond, msg)                          \
  do {                                          \
    if (!(cond)) return fail(GB_EINVAL, msg);   \
  } while (0)

#include "alloc.inc"

extern "C" const char *gb_last_error(void) { return g_err.c_str(); }
extern "C" const char *gb_version(void) { return "genb200 0.1 (sm_100a)"; }

static int require_gpu() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(GB_ENOGPU, "no CUDA device: libgenb200 has no CPU fallback");
  }
  return GB_OK;
}
extern "C" int gb_device_count(int *count) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
  *count = n;
  return GB_OK;
}
extern "C" int gb_set_device(int device) {
  int rc = require_gpu();
  if (rc) return rc;
  CK(cudaSetDevice(device));
  return GB_OK;
}

// ------------------------------------------------------------------------------------------
// models
// ------------------------------------------------------------------------------------------
struct gb_model {
  int kind = 0;
  std::vector<double> params;
  int S = 0, nobs = 1, K = 0, M = 0, D = 0;
  // GB_MODEL_SPEC: parsed program
  std::vector<SpecOp> ops;
  std::vector<double> spec_params;
  int n_init_ops = 0, n_step_ops = 0;
  int nn[2] = {0, 0}, nu[2] = {0, 0};   // normal / uniform draws of the step [0] and generate [1] programs
  long long spec_id = 0;
};
static std::atomic<long long> g_spec_next_id{1};
static long long g_spec_loaded[64] = {0};   // per device: id of the program currently in the interpreter's constant memory
static_assert((int)DIST_UNIFORM == (int)GB_DIST_UNIFORM && (int)DIST_EXPONENTIAL == (int)GB_DIST_EXPONENTIAL &&
              (int)DIST_LAPLACE == (int)GB_DIST_LAPLACE && (int)DIST_CAUCHY == (int)GB_DIST_CAUCHY &&
              (int)DIST_INV_GAMMA == (int)GB_DIST_INV_GAMMA && (int)DIST_BETA == (int)GB_DIST_BETA &&
              (int)DIST_POISSON == (int)GB_DIST_POISSON && (int)DIST_GEOMETRIC == (int)GB_DIST_GEOMETRIC &&
              (int)DIST_BINOM == (int)GB_DIST_BINOM && (int)DIST_NEG_BINOM == (int)GB_DIST_NEG_BINOM &&
              (int)DIST_UNIFORM_DISCRETE == (int)GB_DIST_UNIFORM_DISCRETE && (int)DIST_BETA_UNIFORM == (int)GB_DIST_BETA_UNIFORM &&
              (int)DIST_DIRICHLET == (int)GB_DIST_DIRICHLET, "distribution ids");

static int model_state_dim(int kind, const double *p) {
  switch (kind) {
    case MODEL_LGSSM: case MODEL_SV: case MODEL_HMM: return 1;
    case MODEL_BEARINGS: return 4;
    case MODEL_BLR: return (int)p[0];
    case MODEL_SPEC: return (int)p[0];
  }
  return -1;
}
static int model_num_normals(int kind, const double *p, int is_init) {
  switch (kind) {
    case MODEL_LGSSM: case MODEL_SV: return 1;
    case MODEL_BEARINGS: return is_init ? 4 : 2;
    case MODEL_HMM: return 0;
    case MODEL_BLR: return is_init ? (int)p[0] : 0;
  }
  return 0;
}
static int model_num_uniforms(int kind) { return kind == MODEL_HMM ? 1 : 0; }

static bool host_cholesky(const double *cov, int k, std::vector<double> &L, double *logdet);
extern "C" int gb_model_create(int kind, const double *params, int nparams, gb_model_t **out) {
  REQ(out && params && nparams > 0, "gb_model_create: null argument");
  int need = 0;
  switch (kind) {
    case MODEL_LGSSM: need = 5; break;
    case MODEL_SV: need = 4; break;
    case MODEL_BEARINGS: need = 10; break;
    case MODEL_HMM: {
      REQ(nparams >= 2, "HMM params: [K, M, prior, transition, emission]");
      int K = (int)params[0], M = (int)params[1];
      REQ(K >= 1 && K <= 1024 && M >= 1 && M <= 1024, "HMM: K, M must be in 1..1024");
      need = 2 + K + K * K + M * K;
      break;
    }
    case MODEL_BLR: {
      REQ(nparams >= 3, "BLR params: [D, n, sigma, X]");
      int D = (int)params[0], n = (int)params[1];
      if (!(D >= 1 && D <= kBlrMax && n >= 1 && n <= kBlrMax))
        return fail(GB_EUNSUPPORTED, "BLR: D and n must be in 1..64");
      need = 3 + n * D;
      break;
    }
    case MODEL_SPEC: {
      REQ(nparams >= 5, "SPEC params: [S, n_params, n_obs, n_init_ops, n_step_ops, params..., ops (code, dst, a, b, imm)...]");
      const int S = (int)params[0], np_ = (int)params[1], no = (int)params[2], ni = (int)params[3], ns = (int)params[4];
      if (!(S >= 1 && S <= kSpecMaxState && np_ >= 0 && np_ <= kSpecMaxParams && no >= 1 && no <= 64 && ni >= 1 && ns >= 0 &&
            ni + ns <= kSpecMaxOps))
        return fail(GB_EUNSUPPORTED, "SPEC: limits are 16 state fields, 2048 parameters, 64 observations, 256 ops");
      need = 5 + np_ + 5 * (ni + ns);
      break;
    }
    default:
      return fail(GB_EUNSUPPORTED, "gb_model_create: model kind outside the supported class");
  }
  REQ(nparams == need, "gb_model_create: wrong number of parameters for this model kind");
  gb_model *m = new gb_model;
  m->kind = kind;
  m->params.assign(params, params + nparams);
  if (kind == MODEL_SPEC) {
    const int np_ = (int)params[1], ni = (int)params[3], ns = (int)params[4];
    m->nobs = (int)params[2];
    m->n_init_ops = ni; m->n_step_ops = ns;
    m->spec_params.assign(params + 5, params + 5 + np_);
    const double *q = params + 5 + np_;
    int stored[2][kSpecMaxState] = {{0}};
    const int S = (int)params[0];
    for (int k = 0; k < ni + ns; k++, q += 5) {
      SpecOp op{(int)q[0], (int)q[1], (int)q[2], (int)q[3], q[4]};
      const int prog = k < ni ? 1 : 0;   // 1 = generate, 0 = step
      bool ok = op.dst >= 0 && op.dst < kSpecRegs;
      auto reg = [](int r) { return r >= 0 && r < kSpecRegs; };
      switch (op.code) {
        case OP_CONST: case OP_TIME: break;
        case OP_PREV: ok = ok && op.a >= 0 && op.a < S && prog == 0; break;
        case OP_PARAM: ok = ok && op.a >= 0 && op.a < np_; break;
        case OP_OBS: ok = ok && op.a >= 0 && op.a < m->nobs; break;
        case OP_ADD: case OP_SUB: case OP_MUL: case OP_DIV: case OP_ATAN2: ok = ok && reg(op.a) && reg(op.b); break;
        case OP_NEG: case OP_SQRT: case OP_EXP: case OP_LOG: case OP_ABS: ok = ok && reg(op.a); break;
        case OP_SAMPLE_NORMAL: ok = ok && reg(op.a) && reg(op.b); m->nn[prog]++; break;
        case OP_OBSERVE_NORMAL: case OP_OBSERVE_GAMMA: ok = ok && reg(op.a) && reg(op.b); break;
        case OP_SAMPLE_GAMMA: ok = ok && reg(op.a) && reg(op.b); break;
        case OP_SAMPLE_BERNOULLI: ok = ok && reg(op.a); m->nu[prog]++; break;
        case OP_OBSERVE_BERNOULLI: ok = ok && reg(op.a); break;
        case OP_SAMPLE_CATEGORICAL: ok = ok && reg(op.a) && op.b >= 1 && op.b <= np_; m->nu[prog]++; break;
        case OP_OBSERVE_CATEGORICAL: ok = ok && reg(op.a) && op.b >= 1 && op.b <= np_; break;
        case OP_STORE: ok = op.dst >= 0 && op.dst < S && reg(op.a); if (ok) stored[prog][op.dst]++; break;
        case OP_WEIGHT: ok = reg(op.a); break;
        case OP_LOGPDF_NORMAL: ok = ok && reg(op.a) && reg(op.b) && reg((int)op.imm); break;
        case OP_SAMPLE_DIST: {      // inverse-CDF draw of built-in (int)imm from the uniform stream
          const int di = (int)op.imm;
          ok = ok && reg(op.a) && reg(op.b) && (di == DIST_UNIFORM || di == DIST_EXPONENTIAL || di == DIST_LAPLACE || di == DIST_CAUCHY ||
                                               di == DIST_GEOMETRIC || di == DIST_UNIFORM_DISCRETE);
          m->nu[prog]++;
          break;
        }
        case OP_OBSERVE_DIST: ok = ok && reg(op.a) && reg(op.b) && (int)op.imm >= DIST_UNIFORM && (int)op.imm <= DIST_UNIFORM_DISCRETE; break;
        case OP_LOGPDF_DIST: {
          const int di = (int)op.imm % 100, xr = (int)op.imm / 100;
          ok = ok && reg(op.a) && reg(op.b) && reg(xr) && di >= DIST_UNIFORM && di <= DIST_UNIFORM_DISCRETE;
          break;
        }
        case OP_SAMPLE_MVNORMAL: case OP_OBSERVE_MVNORMAL: {
          const int kk = (int)op.imm;
          ok = kk >= 1 && kk <= kSpecMaxMv && op.dst >= 0 && op.dst + kk <= kSpecRegs && op.a >= 0 && op.a + kk <= kSpecRegs &&
               op.b >= 0 && op.b + kk * kk <= np_;
          if (ok) {
            // mvnormal.jl:14-15 builds MvNormal(mu, Symmetric(cov)) -- a Cholesky factorisation -- at EVERY call; the covariance
            // is a parameter of the model here, so it is factored once: [L (k*k, row-major lower), logdet] appended to the parameters
            std::vector<double> Lf;
            double logdet = 0.0;
            if (!host_cholesky(params + 5 + op.b, kk, Lf, &logdet)) { delete m; return fail(GB_EINVAL, "SPEC: mvnormal covariance of op #" + std::to_string(k) + " is not positive definite"); }
            if ((int)m->spec_params.size() + kk * kk + 1 > kSpecMaxParams) { delete m; return fail(GB_EUNSUPPORTED, "SPEC: parameter block full"); }
            op.b = (int)m->spec_params.size();
            m->spec_params.insert(m->spec_params.end(), Lf.begin(), Lf.end());
            m->spec_params.push_back(logdet);
            if (op.code == OP_SAMPLE_MVNORMAL) m->nn[prog] += kk;
          }
          break;
        }
        default: ok = false;
      }
      if (!ok) { delete m; return fail(GB_EINVAL, "SPEC: malformed op #" + std::to_string(k)); }
      m->ops.push_back(op);
    }
    for (int f = 0; f < S; f++)
      if (stored[1][f] != 1 || (ns > 0 && stored[0][f] != 1)) {
        delete m;
        return fail(GB_EINVAL, "SPEC: every state field must be stored exactly once by the generate program and by the step program");
      }
    m->spec_id = g_spec_next_id++;
  }
  m->S = model_state_dim(kind, params);
  if (kind == MODEL_HMM) { m->K = (int)params[0]; m->M = (int)params[1]; }
  if (kind == MODEL_BLR) { m->D = (int)params[0]; m->nobs = (int)params[1]; }
  *out = m;
  return GB_OK;
}
extern "C" void gb_model_free(gb_model_t *m) { delete m; }
extern "C" int gb_model_state_dim(const gb_model_t *m) { return m ? m->S : -1; }
extern "C" int gb_model_num_obs(const gb_model_t *m) { return m ? m->nobs : -1; }
extern "C" int gb_model_num_normals(const gb_model_t *m, int, int is_init) {
  if (m && m->kind == MODEL_SPEC) return m->nn[is_init ? 1 : 0];
  return m ? model_num_normals(m->kind, m->params.data(), is_init) : -1;
}
extern "C" int gb_model_num_uniforms(const gb_model_t *m, int, int is_init) {
  if (m && m->kind == MODEL_SPEC) return m->nu[is_init ? 1 : 0];
  return m ? model_num_uniforms(m->kind) : -1;
}

// ------------------------------------------------------------------------------------------
// particle store
// ------------------------------------------------------------------------------------------
// what is known about the current log weights
//   W_STATS: h_stats (m, s1, s2, lse, ess, q_total) valid on the host     W_MAX: d_stats->m valid on the device
//   W_ZERO : all log weights are exactly 0.0 (just resampled)             W_DIRTY: nothing known
enum WState { W_STATS = 0, W_ZERO = 1, W_DIRTY = 2, W_MAX = 3 };

struct TimedLaunch { int k; cudaEvent_t a, b; };

struct gb_pf {
  gb_model model;
  int dtype = GB_F64;
  size_t esz = 8;
  long long n = 0, n_global = 0, pid_offset = 0, ld = 0;
  long long n_pad = 0, ext_cap = 0;   // rows [n_pad, n_pad + ext_cap) of every column: offspring received from other shards
  long long *ext_parents = nullptr;   // global 1-based parents of the received rows
  bool ext_used = false;
  uint64_t seed = 0;
  int step = 0;
  int S = 0;
  void *state[2] = {nullptr, nullptr};
  int cur = 0;
  void *logw = nullptr, *incr = nullptr;
  int *anc = nullptr;             // local ancestors of the last resample (n entries, padded)
  bool anc_valid = false;
  bool pending_gather = false;    // resample decided, rows not moved yet (lazy gather)
  bool lazy = true;
  bool mh_ok = false;             // state[cur ^ 1] still holds the parents of the current traces (MH rejuvenation)
  bool mh_parent_anc = false;     // ... reached through `anc` (the last step gathered through it)
  unsigned int mh_moves = 0;
  void *lik = nullptr;            // GB_MODEL_SPEC rejuvenation: score of every particle's constrained choices (kernels_mh.cuh)
  int lik_step = -1;              // step index `lik` belongs to
  int incr_step = -1;             // step index whose log incremental weights `incr` holds
  int incr_where = 0;             // ... and where they are: 0 = `incr`; 1 = `logw` (gathered step: logw was reset to 0, so logw == incr);
                                  // 2 = ask the device (d_stats->incr_in_logw: the step's gather decision was taken there)
  int spec_backend = -1;          // GB_MODEL_SPEC: 1 = NVRTC-specialised kernel, 0 = interpreter (last launch)      // rejuvenation sweeps applied since the last step (Philox draw-index block)
  long long *parents64 = nullptr; // global 1-based parents (sharded path)
  bool parents64_valid = false;
  double log_ml_est = 0.0;
  int wstate = W_DIRTY;
  double2 *partials = nullptr;     // per-CTA partials (running max as double, or (s1, s2))
  // optional history: per step a snapshot of the trace columns and, when a resample preceded the
  // step, its ancestors -- enough to reconstruct every particle's full trajectory (Gen's traces hold all choices)
  bool history = false;
  bool hist_resampled = false;      // a resample happened since the last recorded step
  std::vector<void *> hist_state;   // [t] -> S columns x n_pad of the filter dtype (t = 0: after generate)
  std::vector<int *> hist_anc;      // [t] -> ancestors applied BEFORE step t (nullptr: no resample)
  // optional window of the last W states of every particle's own lineage (sliding-window rejuvenation, kernels_mh.cuh)
  int win_W = 0, win_len = 0, win_cur = 0;
  void *win[2] = {nullptr, nullptr};           // [W][S][ld] each, double-buffered like the traces
  RunCtl *d_ctl = nullptr, *h_ctl = nullptr;   // fused multi-step run (gb_pf_run) / asynchronous sharded run
  ShardPlanDev *d_shplan = nullptr, *h_shplan = nullptr;
  XchgMail *mail = nullptr;         // this shard's mailbox of the peer-memory exchange (kernels_step.cuh), written by peers
  XchgCtl *d_xc = nullptr, *h_xc = nullptr;
  bool xc_on = false;               // d_stats->xc is attached: every publish_max posts this shard's max to its peers
  bool max_posted = false;          // ... and the max of the CURRENT weights has been posted
  int device = 0;                   // the CUDA device that owns every buffer of this handle
  std::vector<void *> ipc_opened;   // peer allocations mapped through CUDA IPC (closed with the handle)
  long long *d_offs = nullptr;
  int a_world = 0, a_rank = 0;
  long long a_cap = 0;
  bool ws_clean = false;           // tile_q / super_q are all zero (ready to accumulate)
  bool tiles_valid = false;        // tile_q / super_q hold the quantised mass of the current weights w.r.t. h_stats->m
  unsigned int *ticket = nullptr;
  WeightStats *d_stats = nullptr, *h_stats = nullptr;
  // validation noise
  double *d_normals = nullptr, *d_uniforms = nullptr;
  int nn_set = 0, nu_set = 0, nn_cap = 0, nu_cap = 0;
  bool use_u0 = false;
  uint32_t u0 = 0;
  uint64_t *d_u64 = nullptr;
  bool u64_set = false;
  // sorted multinomial (kernels_msort.cuh): spacings e[0..n] (padded to whole tiles), slot-tile sums + prefixes, particle-tile
  // CDF, slot-tile -> particle-tile partition, totals
  unsigned long long *ms_e = nullptr, *ms_st = nullptr, *ms_tcdf = nullptr;
  int *ms_pt = nullptr;
  MsortPlan *ms_plan = nullptr;
  bool ms_e_set = false;            // ms_e holds caller-supplied spacings (gb_pf_set_resample_spacings)
  // the Philox spacings depend on (seed, step) only, not on the weights: the raw-array entry generates them on a second
  // stream while the max and statistics passes run (an issue-bound kernel next to two bandwidth-bound ones)
  cudaStream_t ms_aux = nullptr;
  cudaEvent_t ms_fork = nullptr, ms_join = nullptr;
  bool ms_e_async = false;          // ms_e / the slot-tile sums are being produced on ms_aux: wait for ms_join, skip the spacing pass
  // resampling workspace
  long long ntiles = 0;
  uint64_t *tile_q = nullptr;
  unsigned long long *qbuf = nullptr;   // quantised weights of the last statistics pass (ld entries): what the sweeps scan
  void *mid_ws = nullptr;               // persistent run (kernels_mid.cuh): per-CTA partials + the grid barrier counter
  int mid_grid = 0;
  unsigned long long *super_q = nullptr, *super_excl = nullptr;
  long long nsuper = 0;
  OverflowEntry *oflow = nullptr;
  unsigned int *oflow_count = nullptr;
  ResamplePlan *d_plan = nullptr, *h_plan = nullptr;
  uint64_t *cdf = nullptr;
  int *off_anc = nullptr;         // sharded: ancestors of the shard's offspring range
  long long off_cap = 0, off_lo = 0, off_hi = 0;
  int *guide = nullptr;              // multinomial: guide table (systematic ancestors for u0 = 0)
  uint64_t *tile_save = nullptr;     // multinomial: the tile totals while the guide sweep consumes them
  double *d_tab = nullptr, *d_ptab = nullptr;
  int ptab_cap = 0;
  cudaStream_t stream = 0;
  int num_sms = 148;
  // instrumentation
  bool timing = false;
  std::vector<TimedLaunch> pending;
  std::vector<cudaEvent_t> pool;
  double ms[GB_K_COUNT] = {0};
  long long launches[GB_K_COUNT] = {0};
};

static const char *kKernelNames[GB_K_COUNT] = {"step(gather+propose+weight+max)", "weight_stats(lse+ess+qtotals)", "qtotal(residual)",
                                               "tilescan", "ancestors", "gather", "misc"};
extern "C" const char *gb_kernel_name(int k) { return (k >= 0 && k < GB_K_COUNT) ? kKernelNames[k] : "?"; }

static cudaEvent_t get_event(gb_pf *pf) {
  if (!pf->pool.empty()) { cudaEvent_t e = pf->pool.back(); pf->pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
struct LaunchScope {
  gb_pf *pf; int k; cudaEvent_t a = nullptr, b = nullptr;
  LaunchScope(gb_pf *p, int kid) : pf(p), k(kid) {
    pf->launches[k]++;
    if (pf->timing) { a = get_event(pf); b = get_event(pf); cudaEventRecord(a, pf->stream); }
  }
  ~LaunchScope() {
    if (pf->timing) { cudaEventRecord(b, pf->stream); pf->pending.push_back({k, a, b}); }
  }
};
#define LAUNCH(pf, kid, ...)      \
  do {                            \
    LaunchScope ls__(pf, kid);    \
    __VA_ARGS__;                  \
  } while (0)

// Kernels of the per-step chain (weight statistics -> super-tile scan -> ancestors -> overflow -> step) are launched
// with programmatic dependent launch: each starts while its predecessor drains and waits (pdl_wait) before reading.
// GENB200_PDL=0 launches them plainly.
static bool pdl_enabled() {
  static const bool on = [] { const char *e = getenv("GENB200_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
template <typename... P, typename... A>
static cudaError_t klaunch(void (*kern)(P...), unsigned grid, unsigned block, cudaStream_t st, A &&...args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(block, 1, 1);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...);
}

static std::unordered_map<const void *, int> g_occ;
static std::mutex g_occ_mu;
template <typename F> static int grid_for(gb_pf *pf, F kernel, long long work_items, int block) {
  const void *key = (const void *)kernel;
  int occ;
  {
    std::lock_guard<std::mutex> lk(g_occ_mu);
    auto it = g_occ.find(key);
    if (it == g_occ.end()) {
      occ = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, 0) != cudaSuccess || occ < 1) occ = 1;
      g_occ[key] = occ;
    } else occ = it->second;
  }
  long long blocks = (work_items + block - 1) / block;
  long long cap = (long long)pf->num_sms * occ;
  if (blocks < 1) blocks = 1;
  return (int)std::min(blocks, cap);
}

struct DeviceScope {      // make `dev` current for the scope (handles of a multi-GPU filter live on different devices)
  int prev = -1;
  explicit DeviceScope(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev); else prev = -1;
  }
  ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};
static void pf_release(gb_pf *pf) {
  if (!pf) return;
  DeviceScope ds__(pf->device);
  FreeBatch fb__;
  gb_dev_free(pf->mail); gb_dev_free(pf->d_xc); gb_host_free(pf->h_xc); gb_dev_free(pf->win[0]); gb_dev_free(pf->win[1]);
  gb_dev_free(pf->state[0]); gb_dev_free(pf->state[1]); gb_dev_free(pf->logw); gb_dev_free(pf->incr); gb_dev_free(pf->anc);
  gb_dev_free(pf->parents64); gb_dev_free(pf->ext_parents); gb_dev_free(pf->partials); gb_dev_free(pf->ticket); gb_dev_free(pf->d_stats);
  gb_host_free(pf->h_stats); gb_dev_free(pf->d_normals); gb_dev_free(pf->d_uniforms); gb_dev_free(pf->d_u64);
  if (pf->ms_aux) { cudaStreamSynchronize(pf->ms_aux); cudaStreamDestroy(pf->ms_aux); cudaEventDestroy(pf->ms_fork); cudaEventDestroy(pf->ms_join); }
  gb_dev_free(pf->ms_e); gb_dev_free(pf->ms_st); gb_dev_free(pf->ms_tcdf); gb_dev_free(pf->ms_pt); gb_dev_free(pf->ms_plan);
  gb_dev_free(pf->d_ctl); gb_host_free(pf->h_ctl); gb_dev_free(pf->d_shplan); gb_host_free(pf->h_shplan); gb_dev_free(pf->d_offs); gb_dev_free(pf->super_q); gb_dev_free(pf->super_excl); gb_dev_free(pf->oflow); gb_dev_free(pf->oflow_count);
  gb_dev_free(pf->tile_q); gb_dev_free(pf->qbuf); gb_dev_free(pf->d_plan); gb_dev_free(pf->mid_ws);
  gb_dev_free(pf->lik); gb_dev_free(pf->guide); gb_dev_free(pf->tile_save);
  gb_host_free(pf->h_plan); gb_dev_free(pf->cdf); gb_dev_free(pf->off_anc); gb_dev_free(pf->d_tab); gb_dev_free(pf->d_ptab);
  for (auto p : pf->ipc_opened) cudaIpcCloseMemHandle(p);
  for (auto p : pf->hist_state) gb_dev_free(p);
  for (auto p : pf->hist_anc) gb_dev_free(p);
  for (auto &t : pf->pending) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
  for (auto e : pf->pool) cudaEventDestroy(e);
  delete pf;
}
extern "C" void gb_pf_free(gb_pf_t *pf) { pf_release(pf); }

static const int kMaxPartials = 148 * 16 + 64;

// workspace of the systematic pipeline: super-tile totals / prefixes and the overflow queue
static cudaError_t alloc_sys_workspace(gb_pf *pf) {
  pf->nsuper = (pf->ntiles + kSuper - 1) / kSuper;
  cudaError_t e = gb_dev_alloc((void **)&pf->super_q, (size_t)(pf->nsuper + 1) * 8);
  if (e == cudaSuccess) e = gb_dev_alloc((void **)&pf->super_excl, (size_t)(pf->nsuper + 1) * 8);
  if (e == cudaSuccess) e = gb_dev_alloc((void **)&pf->oflow, (size_t)(pf->n_global / kChunk + 2) * sizeof(OverflowEntry));
  if (e == cudaSuccess) e = gb_dev_alloc((void **)&pf->oflow_count, 64);
  if (e == cudaSuccess) e = cudaMemset(pf->oflow_count, 0, 64);
  return e;
}

extern "C" int gb_pf_create(const gb_model_t *m, int64_t n_local, int64_t n_global, int64_t pid_offset, int dtype,
                            uint64_t seed, gb_pf_t **out) {
  REQ(m && out, "gb_pf_create: null argument");
  REQ(n_local >= 1 && n_global >= n_local && pid_offset >= 0 && pid_offset + n_local <= n_global,
      "gb_pf_create: bad particle counts");
  REQ(n_global < ((int64_t)1 << 31), "gb_pf_create: at most 2^31-1 particles");
  REQ(dtype == GB_F64 || dtype == GB_F32, "gb_pf_create: dtype must be GB_F64 or GB_F32");
  int rc = require_gpu();
  if (rc) return rc;
  gb_pf *pf = new gb_pf;
  pf->model = *m;
  pf->dtype = dtype;
  pf->esz = dtype == GB_F64 ? 8 : 4;
  pf->n = n_local; pf->n_global = n_global; pf->pid_offset = pid_offset;
  pf->n_pad = ((n_local + kTile - 1) / kTile) * kTile;
  // sharded filters keep head-room for rows migrating in at a resample (25% of the shard + one tile);
  // a resample that needs more falls back to the staged exchange (gb_pf_import_rows)
  pf->ext_cap = n_local == n_global ? 0 : ((n_local / 4 + kTile - 1) / kTile + 1) * kTile;
  pf->ld = pf->n_pad + pf->ext_cap;
  pf->seed = seed;
  pf->S = m->S;
  pf->ntiles = pf->n_pad / kTile;
  int dev = 0;
  cudaGetDevice(&dev);
  pf->device = dev;
  cudaDeviceGetAttribute(&pf->num_sms, cudaDevAttrMultiProcessorCount, dev);
  auto A = [&](void **p, size_t bytes, bool zero) -> cudaError_t {
    cudaError_t e = gb_dev_alloc(p, bytes);
    if (e == cudaSuccess && zero) e = cudaMemset(*p, 0, bytes);
    return e;
  };
  cudaError_t e = cudaSuccess;
  size_t col = (size_t)pf->ld * pf->esz;
  if (e == cudaSuccess) e = A(&pf->state[0], col * pf->S, true);
  if (e == cudaSuccess) e = A(&pf->state[1], col * pf->S, true);
  if (e == cudaSuccess) e = A(&pf->logw, col, true);
  if (e == cudaSuccess) e = A(&pf->incr, col, true);
  if (e == cudaSuccess) e = A((void **)&pf->anc, (size_t)pf->ld * 4, true);
  if (e == cudaSuccess && pf->ext_cap) e = A((void **)&pf->ext_parents, (size_t)pf->ext_cap * 8, true);
  if (e == cudaSuccess) e = A((void **)&pf->partials, sizeof(double2) * kMaxPartials, true);
  if (e == cudaSuccess) e = A((void **)&pf->ticket, 64, true);
  if (e == cudaSuccess) e = A((void **)&pf->d_stats, sizeof(WeightStats), true);
  if (e == cudaSuccess) e = gb_host_alloc((void **)&pf->h_stats, sizeof(WeightStats));
  if (e == cudaSuccess) e = A((void **)&pf->tile_q, (size_t)(pf->ntiles + 1) * 8, true);
  if (e == cudaSuccess) e = A((void **)&pf->d_plan, sizeof(ResamplePlan), true);
  if (e == cudaSuccess) e = gb_host_alloc((void **)&pf->h_plan, sizeof(ResamplePlan));
  if (e == cudaSuccess) e = alloc_sys_workspace(pf);
  if (e == cudaSuccess && m->kind == MODEL_HMM) {
    size_t nt = m->params.size() - 2;
    e = A((void **)&pf->d_tab, nt * 8, false);
    if (e == cudaSuccess) e = cudaMemcpy(pf->d_tab, m->params.data() + 2, nt * 8, cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) {
    std::string msg = std::string("gb_pf_create: ") + cudaGetErrorString(e);
    pf_release(pf);
    cudaGetLastError();
    return fail(GB_ECUDA, msg);
  }
  *out = pf;
  return GB_OK;
}

// Work enqueued so far (creation-time fills on the stream the handle had, e.g. the legacy default stream) must not be
// overtaken by kernels on the new stream -- a non-blocking stream does not order itself after the default one, and a handle
// built from recycled blocks holds stale bytes (an old exchange pointer in d_stats) until its memsets have run.
extern "C" int gb_pf_set_stream(gb_pf_t *pf, void *s) {
  REQ(pf, "null handle");
  DeviceScope ds__(pf->device);
  CK(cudaStreamSynchronize(pf->stream));
  CK(cudaDeviceSynchronize());
  pf->stream = (cudaStream_t)s;
  return GB_OK;
}
extern "C" int gb_pf_sync(gb_pf_t *pf) { REQ(pf, "null handle"); CK(cudaStreamSynchronize(pf->stream)); return GB_OK; }
extern "C" int gb_pf_set_lazy_gather(gb_pf_t *pf, int lazy) { REQ(pf, "null handle"); pf->lazy = lazy != 0; return GB_OK; }
extern "C" int gb_pf_num_particles(const gb_pf_t *pf, int64_t *nl, int64_t *ng) {
  REQ(pf, "null handle");
  if (nl) *nl = pf->n;
  if (ng) *ng = pf->n_global;
  return GB_OK;
}
extern "C" int gb_pf_step_index(const gb_pf_t *pf, int *step) { REQ(pf && step, "null argument"); *step = pf->step; return GB_OK; }

extern "C" int gb_pf_enable_timing(gb_pf_t *pf, int on) { REQ(pf, "null handle"); pf->timing = on != 0; return GB_OK; }
extern "C" int gb_pf_get_timing(gb_pf_t *pf, double *ms, int64_t *launches) {
  REQ(pf, "null handle");
  CK(cudaStreamSynchronize(pf->stream));
  for (auto &t : pf->pending) {
    float f = 0.f;
    cudaEventElapsedTime(&f, t.a, t.b);
    pf->ms[t.k] += f;
    pf->pool.push_back(t.a);
    pf->pool.push_back(t.b);
  }
  pf->pending.clear();
  for (int k = 0; k < GB_K_COUNT; k++) {
    if (ms) ms[k] = pf->ms[k];
    if (launches) launches[k] = pf->launches[k];
    pf->ms[k] = 0;
    pf->launches[k] = 0;
  }
  return GB_OK;
}

// upload [nd][n] host doubles into a [nd][ld] device buffer
static int upload_rows(gb_pf *pf, double **dbuf, int *cap, const double *src, int nd) {
  if (nd > *cap) {
    gb_dev_free(*dbuf);
    *dbuf = nullptr;
    CK(gb_dev_alloc((void **)dbuf, (size_t)nd * pf->ld * 8));
    CK(cudaMemsetAsync(*dbuf, 0, (size_t)nd * pf->ld * 8, pf->stream));
    *cap = nd;
  }
  CK(cudaMemcpy2DAsync(*dbuf, (size_t)pf->ld * 8, src, (size_t)pf->n * 8, (size_t)pf->n * 8, nd, cudaMemcpyHostToDevice,
                       pf->stream));
  CK(cudaStreamSynchronize(pf->stream));   // the host buffer is pageable and caller-owned
  return GB_OK;
}
extern "C" int gb_pf_set_noise(gb_pf_t *pf, const double *normals, int nn, const double *uniforms, int nu) {
  REQ(pf, "null handle");
  pf->nn_set = pf->nu_set = 0;
  if (normals && nn > 0) {
    int rc = upload_rows(pf, &pf->d_normals, &pf->nn_cap, normals, nn);
    if (rc) return rc;
    pf->nn_set = nn;
  }
  if (uniforms && nu > 0) {
    int rc = upload_rows(pf, &pf->d_uniforms, &pf->nu_cap, uniforms, nu);
    if (rc) return rc;
    pf->nu_set = nu;
  }
  return GB_OK;
}
extern "C" int gb_pf_set_resample_noise(gb_pf_t *pf, int use_u0, uint32_t u0, const uint64_t *u64s) {
  REQ(pf, "null handle");
  pf->use_u0 = use_u0 != 0;
  pf->u0 = u0;
  pf->u64_set = false;
  if (u64s) {
    if (!pf->d_u64) CK(gb_dev_alloc((void **)&pf->d_u64, (size_t)pf->ld * 8));
    CK(cudaMemcpyAsync(pf->d_u64, u64s, (size_t)pf->n * 8, cudaMemcpyHostToDevice, pf->stream));
    CK(cudaStreamSynchronize(pf->stream));
    pf->u64_set = true;
  }
  return GB_OK;
}

static long long msort_slot_tiles(const gb_pf *pf) { return (pf->n + 1 + kTile - 1) / kTile; }   // tiles covering e[0..n]
static int msort_workspace(gb_pf *pf) {
  const long long nst = msort_slot_tiles(pf);
  if (!pf->ms_e) CK(gb_dev_alloc((void **)&pf->ms_e, (size_t)nst * kTile * 8));
  if (!pf->ms_st) CK(gb_dev_alloc((void **)&pf->ms_st, (size_t)(2 * nst + 4) * 8));
  if (!pf->ms_tcdf) CK(gb_dev_alloc((void **)&pf->ms_tcdf, (size_t)(pf->ntiles + 2) * 8));
  if (!pf->ms_pt) CK(gb_dev_alloc((void **)&pf->ms_pt, (size_t)(nst + 2) * 4));
  if (!pf->ms_plan) CK(gb_dev_alloc((void **)&pf->ms_plan, sizeof(MsortPlan)));
  return GB_OK;
}
// spacing pass of GB_RESAMPLE_MULTINOMIAL_SORTED on the auxiliary stream, ordered after everything enqueued on pf->stream so far
static int msort_prelaunch_spacings(gb_pf *pf, uint64_t seed, uint32_t step) {
  int rc = msort_workspace(pf);
  if (rc) return rc;
  if (!pf->ms_aux) {
    CK(cudaStreamCreateWithFlags(&pf->ms_aux, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&pf->ms_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&pf->ms_join, cudaEventDisableTiming));
  }
  const long long nst = msort_slot_tiles(pf);
  CK(cudaEventRecord(pf->ms_fork, pf->stream));
  CK(cudaStreamWaitEvent(pf->ms_aux, pf->ms_fork, 0));
  const int grid = (int)std::min<long long>(nst, (long long)pf->num_sms * 8);
  msort_spacing_kernel<true><<<grid, kBlock, 0, pf->ms_aux>>>(pf->ms_e, pf->n, nst, seed, step, 0ll, pf->ms_st);
  CK(cudaGetLastError());
  CK(cudaEventRecord(pf->ms_join, pf->ms_aux));
  pf->ms_e_async = true;
  return GB_OK;
}
// validation mode of GB_RESAMPLE_MULTINOMIAL_SORTED: the n + 1 integer spacings (each >= 1, sum < 2^63) in place of Philox
extern "C" int gb_pf_set_resample_spacings(gb_pf_t *pf, const uint64_t *spacings) {
  REQ(pf, "null handle");
  pf->ms_e_set = false;
  if (!spacings) return GB_OK;
  int rc = msort_workspace(pf);
  if (rc) return rc;
  CK(cudaMemcpyAsync(pf->ms_e, spacings, (size_t)(pf->n + 1) * 8, cudaMemcpyHostToDevice, pf->stream));
  CK(cudaStreamSynchronize(pf->stream));
  pf->ms_e_set = true;
  return GB_OK;
}

// ------------------------------------------------------------------------------------------
// step / generate dispatch
// ------------------------------------------------------------------------------------------
template <typename R, int KIND, int PK, bool INIT>
static int launch_step_t(gb_pf *pf, const StepArgs &a, bool predrawn, int gather /* 0, 1, 2 = device flag */) {
  constexpr int V = StepV<R>::V;
  const long long nvec = (pf->n + V - 1) / V;
#define GO(PD, GA)                                                                         \
  {                                                                                        \
    auto kern = step_kernel<R, KIND, PK, INIT, PD, GA>;                                    \
    int grid = grid_for(pf, kern, nvec, kBlock);                                           \
    LAUNCH(pf, GB_K_STEP, klaunch(kern, grid, kBlock, pf->stream, a));                     \
  }
  if (INIT) {
    if (predrawn) GO(true, 0) else GO(false, 0)
  } else {
    if (predrawn) { if (gather) GO(true, 1) else GO(true, 0) }
    else { if (gather == 2) GO(false, 2) else if (gather) GO(false, 1) else GO(false, 0) }
  }
#undef GO
  CK(cudaGetLastError());
  return GB_OK;
}
template <typename R, bool INIT>
static int launch_step_k(gb_pf *pf, const StepArgs &a, int pk, bool predrawn, int gather) {
  switch (pf->model.kind) {
    case MODEL_LGSSM:
      if (pk == PROPOSAL_AFFINE_NORMAL) return launch_step_t<R, MODEL_LGSSM, PROPOSAL_AFFINE_NORMAL, INIT>(pf, a, predrawn, gather);
      if (pk == PROPOSAL_DEFAULT) return launch_step_t<R, MODEL_LGSSM, PROPOSAL_DEFAULT, INIT>(pf, a, predrawn, gather);
      break;
    case MODEL_SV:
      if (pk == PROPOSAL_DEFAULT) return launch_step_t<R, MODEL_SV, PROPOSAL_DEFAULT, INIT>(pf, a, predrawn, gather);
      break;
    case MODEL_BEARINGS:
      if (pk == PROPOSAL_DEFAULT) return launch_step_t<R, MODEL_BEARINGS, PROPOSAL_DEFAULT, INIT>(pf, a, predrawn, gather);
      break;
    case MODEL_HMM:
      if (pk == PROPOSAL_CATEGORICAL_TABLE) return launch_step_t<R, MODEL_HMM, PROPOSAL_CATEGORICAL_TABLE, INIT>(pf, a, predrawn, gather);
      if (pk == PROPOSAL_DEFAULT) return launch_step_t<R, MODEL_HMM, PROPOSAL_DEFAULT, INIT>(pf, a, predrawn, gather);
      break;
  }
  return fail(GB_EUNSUPPORTED, "this proposal kind is not supported for this model kind");
}

static int fill_model_params(gb_pf *pf, ModelParams &mp, const double *obs, int nobs, int pk, const double *pargs,
                             int npargs, bool init) {
  memset(&mp, 0, sizeof(mp));
  const gb_model &m = pf->model;
  REQ(obs && nobs >= 1, "observation vector is required");
  mp.obs0 = obs[0];
  mp.K = m.K; mp.M = m.M;
  if (m.kind == MODEL_HMM) {
    mp.tab = pf->d_tab;
    REQ(obs[0] >= 1 && obs[0] <= m.M, "HMM observation outside 1..M");
  } else {
    for (size_t i = 0; i < m.params.size() && i < 12; i++) mp.p[i] = m.params[i];
  }
  if (pk == PROPOSAL_AFFINE_NORMAL) {
    REQ(m.kind == MODEL_LGSSM, "affine-normal proposal is defined for the LGSSM model only");
    REQ(pargs && npargs == 3, "affine-normal proposal takes pargs = [c0, c1, std]");
    for (int i = 0; i < 3; i++) mp.pa[i] = pargs[i];
  } else if (pk == PROPOSAL_CATEGORICAL_TABLE) {
    REQ(m.kind == MODEL_HMM, "categorical-table proposal is defined for the HMM model only");
    int need = init ? m.K : m.K * m.K;
    REQ(pargs && npargs == need, "categorical-table proposal: K probabilities at t=0, K*K afterwards");
    if (pf->ptab_cap < need) {
      gb_dev_free(pf->d_ptab);
      pf->d_ptab = nullptr;
      CK(gb_dev_alloc((void **)&pf->d_ptab, (size_t)m.K * m.K * 8));
      pf->ptab_cap = m.K * m.K;
    }
    CK(cudaMemcpyAsync(pf->d_ptab, pargs, (size_t)need * 8, cudaMemcpyHostToDevice, pf->stream));
    CK(cudaStreamSynchronize(pf->stream));
    mp.ptab = pf->d_ptab;
  } else {
    REQ(pk == PROPOSAL_DEFAULT, "unknown proposal kind");
  }
  return GB_OK;
}

static int check_noise(gb_pf *pf, bool init, bool *predrawn) {
  const gb_model &m = pf->model;
  int nn = model_num_normals(m.kind, m.params.data(), init), nu = model_num_uniforms(m.kind);
  if (m.kind == MODEL_SPEC) { nn = m.nn[init ? 1 : 0]; nu = m.nu[init ? 1 : 0]; }
  *predrawn = (pf->nn_set > 0 || pf->nu_set > 0);
  if (*predrawn) {
    REQ(pf->nn_set >= nn && pf->nu_set >= nu, "gb_pf_set_noise: fewer pre-drawn rows than this call consumes");
  }
  return GB_OK;
}

static int blr_generate(gb_pf *pf, const double *obs, int nobs, bool predrawn);

// GB_MODEL_SPEC: program + parameters into constant memory (once per model), observations per call
#include "spec_jit.inc"

template <typename R, bool INIT>
static int launch_spec(gb_pf *pf, const double *obs, int nobs, bool predrawn, int gather) {
  const gb_model &m = pf->model;
  SpecArgs a;
  memset(&a, 0, sizeof(a));
  a.state_in = INIT ? nullptr : pf->state[pf->cur];
  a.state_out = pf->state[pf->cur ^ 1];
  a.logw = pf->logw; a.incr = pf->incr;
  a.anc = pf->anc; a.ctl = pf->d_ctl;
  a.normals = pf->d_normals; a.uniforms = pf->d_uniforms;
  a.n = pf->n; a.ld = pf->ld; a.pid_offset = pf->pid_offset;
  a.seed = pf->seed; a.step = INIT ? 0u : (unsigned)(pf->step + 1);
  a.op_begin = INIT ? 0 : m.n_init_ops;
  a.op_end = INIT ? m.n_init_ops : m.n_init_ops + m.n_step_ops;
  a.partials = (double *)pf->partials; a.ticket = pf->ticket; a.stats = pf->d_stats;
  for (int k = 0; k < nobs && k < kSpecMaxObs; k++) a.obs[k] = obs[k];
  // the NVRTC-specialised kernel when it can be built, else the interpreter
  if (JitModule *jm = spec_jit_get(m, pf->dtype)) {
    const int v = jit_variant(INIT, predrawn, gather);
    long long blocks = (pf->n + kBlock - 1) / kBlock;
    const int grid = (int)std::max(1ll, std::min(blocks, (long long)pf->num_sms * jm->occ[v]));
    StepArgs sa;   // fast form: the program is a Transition<> of the step kernel itself
    if (jm->fast) {
      memset(&sa, 0, sizeof(sa));
      sa.state_in = a.state_in; sa.state_out = a.state_out; sa.logw = a.logw; sa.incr = a.incr; sa.anc = a.anc; sa.ctl = a.ctl;
      sa.normals = a.normals; sa.uniforms = a.uniforms; sa.n = a.n; sa.ld = a.ld; sa.pid_offset = a.pid_offset;
      sa.seed = a.seed; sa.step = a.step; sa.partials = a.partials; sa.ticket = a.ticket; sa.stats = a.stats;
      sa.mp.obs0 = obs[0];
      for (int k = 1; k < nobs && k < 16; k++) sa.mp.obsx[k - 1] = obs[k];
      sa.mp.step = a.step;
    }
    void *params[] = {jm->fast ? (void *)&sa : (void *)&a};
    CUresult cr = CUDA_SUCCESS;
    CUlaunchConfig lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDimX = grid; lc.gridDimY = lc.gridDimZ = 1;
    lc.blockDimX = kBlock; lc.blockDimY = lc.blockDimZ = 1;
    lc.hStream = (CUstream)pf->stream;
    CUlaunchAttribute lat[1];
    lat[0].id = CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION;
    lat[0].value.programmaticStreamSerializationAllowed = 1;
    lc.attrs = lat;
    lc.numAttrs = pdl_enabled() ? 1 : 0;
    LAUNCH(pf, GB_K_STEP, cr = g_jit.LaunchKernelEx(&lc, jm->fn[v], params, nullptr));
    if (cr != CUDA_SUCCESS) {
      const char *es = nullptr;
      g_jit.GetErrorString(cr, &es);
      return fail(GB_ECUDA, std::string("cuLaunchKernel(spec_jit_kernel): ") + (es ? es : "?"));
    }
    pf->spec_backend = 1;
    return GB_OK;
  }
  pf->spec_backend = 0;
  int spec_dev = 0;
  cudaGetDevice(&spec_dev);
  spec_dev = spec_dev >= 0 && spec_dev < 64 ? spec_dev : 0;
  static std::mutex spec_mu;                          // (the interpreter's program lives in ONE __constant__ array per device)
  std::lock_guard<std::mutex> spec_lk(spec_mu);
  if (g_spec_loaded[spec_dev] != m.spec_id) {
    CK(cudaDeviceSynchronize());                      // another handle's kernels may still be reading the previous program
    CK(cudaMemcpyToSymbolAsync(c_spec_ops, m.ops.data(), m.ops.size() * sizeof(SpecOp), 0, cudaMemcpyHostToDevice, pf->stream));
    if (!m.spec_params.empty())
      CK(cudaMemcpyToSymbolAsync(c_spec_params, m.spec_params.data(), m.spec_params.size() * 8, 0, cudaMemcpyHostToDevice, pf->stream));
    CK(cudaStreamSynchronize(pf->stream));
    g_spec_loaded[spec_dev] = m.spec_id;
  }
#define GO(PD, GA)                                                                         \
  {                                                                                        \
    auto kern = spec_kernel<R, INIT, PD, GA>;                                              \
    int grid = grid_for(pf, kern, pf->n, kBlock);                                          \
    LAUNCH(pf, GB_K_STEP, klaunch(kern, grid, kBlock, pf->stream, a));                     \
  }
  if (INIT) { if (predrawn) GO(true, 0) else GO(false, 0) }
  else if (predrawn) { if (gather) GO(true, 1) else GO(true, 0) }
  else { if (gather == 2) GO(false, 2) else if (gather) GO(false, 1) else GO(false, 0) }
#undef GO
  CK(cudaGetLastError());
  return GB_OK;
}

// introspection of the run-time specialisation (tests, docs): generated source, compile check, backend in use
extern "C" int gb_spec_jit_source(const gb_model_t *m, char *buf, int64_t cap, int64_t *needed) {
  REQ(m && m->kind == MODEL_SPEC, "gb_spec_jit_source: not a GB_MODEL_SPEC model");
  const std::string src = spec_codegen(*m);
  if (needed) *needed = (int64_t)src.size() + 1;
  if (buf && cap > 0) {
    const size_t k = std::min((size_t)cap - 1, src.size());
    memcpy(buf, src.data(), k);
    buf[k] = 0;
  }
  return GB_OK;
}
// compile (not load) the model's kernels for sm_100a: needs libnvrtc only, no GPU.  cubin_bytes may be NULL.
extern "C" int gb_spec_jit_compile(const gb_model_t *m, int dtype, int64_t *cubin_bytes) {
  REQ(m && m->kind == MODEL_SPEC, "gb_spec_jit_compile: not a GB_MODEL_SPEC model");
  REQ(dtype == GB_F64 || dtype == GB_F32, "gb_spec_jit_compile: dtype");
  std::vector<char> cubin;
  std::vector<std::string> lowered;
  std::string log;
  const int rc = spec_jit_compile(*m, dtype, cubin, lowered, log);
  if (rc) return fail(rc, "gb_spec_jit_compile: " + log);
  if (cubin_bytes) *cubin_bytes = (int64_t)cubin.size();
  return GB_OK;
}
// 1: the last generate/step of this filter ran the NVRTC-specialised kernel; 0: the interpreter; -1: not a SPEC model / not run yet
extern "C" int gb_pf_spec_backend(const gb_pf_t *pf, int *backend) {
  REQ(pf && backend, "null argument");
  *backend = pf->model.kind == MODEL_SPEC ? pf->spec_backend : -1;
  return GB_OK;
}

// One `generate` (init) or extend-by-one `update` sweep over the shard for whatever model kind the handle holds.
// gather: 0 direct input, 1 through the pending ancestors, 2 decided by the device (RunCtl).
static int launch_model(gb_pf *pf, const double *obs, int nobs, int pk, const double *pargs, int npargs, bool init,
                        bool predrawn, int gather) {
  if (pf->model.kind == MODEL_SPEC) {
    REQ(pk == PROPOSAL_DEFAULT, "SPEC models carry their proposal inside the program (default proposal only)");
    REQ(init || pf->model.n_step_ops > 0, "this SPEC model has no step program");
    if (!init) pf->incr_where = 0;      // (the SPEC kernels always store the increments in their own column)
    if (pf->dtype == GB_F64) return init ? launch_spec<double, true>(pf, obs, nobs, predrawn, 0) : launch_spec<double, false>(pf, obs, nobs, predrawn, gather);
    return init ? launch_spec<float, true>(pf, obs, nobs, predrawn, 0) : launch_spec<float, false>(pf, obs, nobs, predrawn, gather);
  }
  StepArgs a;
  memset(&a, 0, sizeof(a));
  int rc = fill_model_params(pf, a.mp, obs, nobs, pk, pargs, npargs, init);
  if (rc) return rc;
  a.state_in = init ? nullptr : pf->state[pf->cur];
  a.state_out = pf->state[pf->cur ^ 1];
  a.logw = pf->logw; a.incr = pf->incr;
  a.anc = pf->anc; a.ctl = pf->d_ctl;
  a.normals = pf->d_normals; a.uniforms = pf->d_uniforms;
  a.n = pf->n; a.ld = pf->ld; a.pid_offset = pf->pid_offset;
  a.seed = pf->seed; a.step = init ? 0u : (unsigned)(pf->step + 1);
  a.partials = (double *)pf->partials; a.ticket = pf->ticket; a.stats = pf->d_stats;
  if (init) return pf->dtype == GB_F64 ? launch_step_k<double, true>(pf, a, pk, predrawn, 0) : launch_step_k<float, true>(pf, a, pk, predrawn, 0);
  a.share_incr = 1;
  rc = pf->dtype == GB_F64 ? launch_step_k<double, false>(pf, a, pk, predrawn, gather) : launch_step_k<float, false>(pf, a, pk, predrawn, gather);
  if (!rc) pf->incr_where = gather;      // 0: own column; 1: shared with logw; 2: decided on the device (d_stats->incr_in_logw)
  return rc;
}

// snapshot of the current trace columns (+ the ancestors that were applied on the way into this step)
static int history_record(gb_pf *pf, bool resampled) {
  void *snap = nullptr;
  const size_t bytes = (size_t)pf->S * pf->n_pad * pf->esz;
  CK(gb_dev_alloc(&snap, bytes));
  CK(cudaMemcpy2DAsync(snap, (size_t)pf->n_pad * pf->esz, pf->state[pf->cur], (size_t)pf->ld * pf->esz, (size_t)pf->n_pad * pf->esz, pf->S,
                       cudaMemcpyDeviceToDevice, pf->stream));
  int *anc = nullptr;
  if (resampled) {
    CK(gb_dev_alloc((void **)&anc, (size_t)pf->n_pad * 4));
    CK(cudaMemcpyAsync(anc, pf->anc, (size_t)pf->n_pad * 4, cudaMemcpyDeviceToDevice, pf->stream));
  }
  pf->hist_state.push_back(snap);
  pf->hist_anc.push_back(anc);
  return GB_OK;
}
extern "C" int gb_pf_enable_history(gb_pf_t *pf, int on) {
  REQ(pf, "null handle");
  REQ(pf->n == pf->n_global, "history is kept for single-shard filters only");
  REQ(!(on && pf->win_W > 0), "gb_pf_enable_history: not together with gb_pf_enable_window");
  pf->history = on != 0;
  return GB_OK;
}
extern "C" int gb_pf_history_length(const gb_pf_t *pf, int *steps) {
  REQ(pf && steps, "null argument");
  *steps = (int)pf->hist_state.size();
  return GB_OK;
}
// out[t][i] = value of trace column `field` at time t on the lineage of the CURRENT particle i (t = 0..len-1)
extern "C" int gb_pf_get_trajectory(gb_pf_t *pf, int field, double *out) {
  REQ(pf && out, "null argument");
  REQ(field >= 0 && field < pf->S, "gb_pf_get_trajectory: field out of range");
  const int len = (int)pf->hist_state.size();
  REQ(pf->history && len > 0, "gb_pf_get_trajectory: history is off (gb_pf_enable_history before generate)");
  REQ(len == pf->step + 1, "gb_pf_get_trajectory: history does not cover every step (gb_pf_run does not record history)");
  int *idx = nullptr;
  double *d_out = nullptr;
  CK(gb_dev_alloc((void **)&idx, (size_t)pf->n_pad * 4));
  cudaError_t e = gb_dev_alloc((void **)&d_out, (size_t)pf->n * 8);
  if (e != cudaSuccess) { gb_dev_free(idx); return fail(GB_ECUDA, cudaGetErrorString(e)); }
  int grid = grid_for(pf, iota_int_kernel, pf->n, kBlock);
  // a resample since the last recorded step (deferred or already gathered) is not in the last snapshot:
  // start from its ancestors
  if (pf->hist_resampled) e = cudaMemcpyAsync(idx, pf->anc, (size_t)pf->n * 4, cudaMemcpyDeviceToDevice, pf->stream);
  else iota_int_kernel<<<grid, kBlock, 0, pf->stream>>>(idx, pf->n);
  for (int t = len - 1; t >= 0 && e == cudaSuccess; t--) {
    const char *snap = (const char *)pf->hist_state[t] + (size_t)field * pf->n_pad * pf->esz;
    if (pf->dtype == GB_F64) backtrace_kernel<double><<<grid, kBlock, 0, pf->stream>>>((const double *)snap, pf->hist_anc[t], idx, pf->n, d_out);
    else backtrace_kernel<float><<<grid, kBlock, 0, pf->stream>>>((const float *)snap, pf->hist_anc[t], idx, pf->n, d_out);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out + (size_t)t * pf->n, d_out, (size_t)pf->n * 8, cudaMemcpyDeviceToHost, pf->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(pf->stream);
  }
  gb_dev_free(idx); gb_dev_free(d_out);
  if (e != cudaSuccess) return fail(GB_ECUDA, std::string("gb_pf_get_trajectory: ") + cudaGetErrorString(e));
  return GB_OK;
}

// ---- window of the last W states per particle (sliding-window rejuvenation) ----
extern "C" int gb_pf_enable_window(gb_pf_t *pf, int W) {
  REQ(pf, "null handle");
  REQ(W >= 0 && W <= kWinMax, "gb_pf_enable_window: 0 <= W <= 8");
  REQ(pf->n == pf->n_global, "gb_pf_enable_window: single-shard filters only");
  REQ(!(W > 0 && pf->history), "gb_pf_enable_window: not together with gb_pf_enable_history (window moves rewrite the recent past)");
  DeviceScope ds__(pf->device);
  {
    FreeBatch fb__;
    gb_dev_free(pf->win[0]); gb_dev_free(pf->win[1]);
  }
  pf->win[0] = pf->win[1] = nullptr;
  pf->win_W = W; pf->win_len = 0; pf->win_cur = 0;
  if (W > 0) {
    const size_t bytes = (size_t)W * pf->S * pf->ld * pf->esz;
    CK(gb_dev_alloc(&pf->win[0], bytes));
    CK(gb_dev_alloc(&pf->win[1], bytes));
    CK(cudaMemsetAsync(pf->win[0], 0, bytes, pf->stream));
    CK(cudaMemsetAsync(pf->win[1], 0, bytes, pf->stream));
  }
  return GB_OK;
}
// shift = 1: after particle_filter_step! (parent = the buffer the step read, through `anc` iff it gathered);
// shift = 0: the eager resample gather moves the window with the particles
static int window_update(gb_pf *pf, const void *parent_state, const int *anc, int shift) {
  if (pf->win_W == 0) return GB_OK;
  if (pf->dtype == GB_F64) {
    auto kern = window_shift_kernel<double>;
    int grid = grid_for(pf, kern, pf->n, kBlock);
    LAUNCH(pf, GB_K_GATHER, kern<<<grid, kBlock, 0, pf->stream>>>((const double *)parent_state, (const double *)pf->win[pf->win_cur],
                                                                   (double *)pf->win[pf->win_cur ^ 1], anc, pf->S, pf->win_W, shift, pf->n, pf->ld));
  } else {
    auto kern = window_shift_kernel<float>;
    int grid = grid_for(pf, kern, pf->n, kBlock);
    LAUNCH(pf, GB_K_GATHER, kern<<<grid, kBlock, 0, pf->stream>>>((const float *)parent_state, (const float *)pf->win[pf->win_cur],
                                                                   (float *)pf->win[pf->win_cur ^ 1], anc, pf->S, pf->win_W, shift, pf->n, pf->ld));
  }
  CK(cudaGetLastError());
  pf->win_cur ^= 1;
  if (shift) pf->win_len = std::min(pf->win_len + 1, pf->win_W);
  return GB_OK;
}

extern "C" int gb_pf_generate(gb_pf_t *pf, const double *obs, int nobs, int pk, const double *pargs, int npargs) {
  REQ(pf, "null handle");
  REQ(nobs == pf->model.nobs, "gb_pf_generate: wrong number of observations");
  bool predrawn;
  int rc = check_noise(pf, true, &predrawn);
  if (rc) return rc;
  pf->step = 0;
  pf->win_len = 0;
  pf->log_ml_est = 0.0;               // particle_filter.jl:133,153
  pf->anc_valid = pf->parents64_valid = pf->pending_gather = false;   // parents = collect(1:N)
  if (pf->model.kind == MODEL_BLR) {
    REQ(pk == PROPOSAL_DEFAULT, "BLR supports the default proposal only");
    rc = blr_generate(pf, obs, nobs, predrawn);
  } else {
    rc = launch_model(pf, obs, nobs, pk, pargs, npargs, true, predrawn, 0);
  }
  if (rc) return rc;
  pf->cur ^= 1;
  pf->wstate = W_MAX;
  pf->max_posted = pf->xc_on;
  pf->tiles_valid = false;
  pf->nn_set = pf->nu_set = 0;
  pf->mh_ok = pf->model.kind != MODEL_BLR; pf->mh_parent_anc = false; pf->mh_moves = 0; pf->lik_step = -1;
  if (pf->history) {
    for (auto p : pf->hist_state) gb_dev_free(p);
    for (auto p : pf->hist_anc) gb_dev_free(p);
    pf->hist_state.clear(); pf->hist_anc.clear();
    if ((rc = history_record(pf, false))) return rc;
  }
  return GB_OK;
}

extern "C" int gb_pf_initialize(const gb_model_t *m, int64_t n, int dtype, uint64_t seed, const double *obs, int nobs,
                                int pk, const double *pargs, int npargs, gb_pf_t **out) {
  gb_pf_t *pf = nullptr;
  int rc = gb_pf_create(m, n, n, 0, dtype, seed, &pf);
  if (rc) return rc;
  rc = gb_pf_generate(pf, obs, nobs, pk, pargs, npargs);
  if (rc) { pf_release(pf); return rc; }
  *out = pf;
  return GB_OK;
}

static int copy_out(gb_pf *pf, const void *dsrc, double *out, long long n);
// where the log incremental weights of the last step live (see gb_pf::incr_where)
static int incr_source(gb_pf *pf, const void **src) {
  if (pf->incr_where == 2) {
    int flag = 0;
    CK(cudaMemcpyAsync(&flag, &pf->d_stats->incr_in_logw, sizeof(int), cudaMemcpyDeviceToHost, pf->stream));
    CK(cudaStreamSynchronize(pf->stream));
    pf->incr_where = flag ? 1 : 0;
  }
  *src = pf->incr_where == 1 ? pf->logw : pf->incr;
  return GB_OK;
}
// call before anything other than a step overwrites logw: the shared copy moves to its own column
static int unshare_incr(gb_pf *pf) {
  if (pf->incr_where == 0 || pf->step < 1 || pf->incr_step != pf->step) { pf->incr_where = 0; return GB_OK; }
  const void *src = nullptr;
  int rc = incr_source(pf, &src);
  if (rc) return rc;
  if (src != pf->incr) CK(cudaMemcpyAsync(pf->incr, pf->logw, (size_t)pf->ld * pf->esz, cudaMemcpyDeviceToDevice, pf->stream));
  pf->incr_where = 0;
  return GB_OK;
}
extern "C" int gb_pf_step(gb_pf_t *pf, const double *obs, int nobs, int pk, const double *pargs, int npargs,
                          double *incr_out) {
  REQ(pf, "null handle");
  REQ(pf->model.kind != MODEL_BLR, "the BLR model has no sequential structure (importance sampling only)");
  REQ(nobs == pf->model.nobs, "gb_pf_step: wrong number of observations");
  bool predrawn;
  int rc = check_noise(pf, false, &predrawn);
  if (rc) return rc;
  const bool gather = pf->pending_gather;
  rc = launch_model(pf, obs, nobs, pk, pargs, npargs, false, predrawn, gather ? 1 : 0);
  if (rc) return rc;
  if ((rc = window_update(pf, pf->state[pf->cur], gather ? pf->anc : nullptr, 1))) return rc;
  pf->step += 1;
  pf->cur ^= 1;                       // particle_filter.jl:203-205, :234-237 (swap traces/new_traces)
  pf->pending_gather = false;
  pf->wstate = W_MAX;
  pf->max_posted = pf->xc_on;
  pf->tiles_valid = false;
  pf->nn_set = pf->nu_set = 0;
  pf->mh_ok = true; pf->mh_parent_anc = gather; pf->mh_moves = 0;
  pf->incr_step = pf->step;
  if (pf->history && (rc = history_record(pf, pf->hist_resampled))) return rc;
  pf->hist_resampled = false;
  if (incr_out) {
    const void *src = nullptr;
    if ((rc = incr_source(pf, &src))) return rc;
    return copy_out(pf, src, incr_out, pf->n);   // staged: pinned chunks + host fan-out (the caller's pages are pageable)
  }
  return GB_OK;
}

// ------------------------------------------------------------------------------------------
// MH rejuvenation of every particle's latest latent choices (kernels_mh.cuh; mh.jl:16-27, :37-58)
// ------------------------------------------------------------------------------------------
template <typename R, int KIND, bool INIT>
static int launch_mh_t(gb_pf *pf, const MhArgs &a, int move, bool predrawn) {
#define GO(MV, PD)                                                                  \
  {                                                                                 \
    auto kern = mh_kernel<R, KIND, INIT, MV, PD>;                                   \
    int grid = grid_for(pf, kern, pf->n, kBlock);                                   \
    LAUNCH(pf, GB_K_MISC, kern<<<grid, kBlock, 0, pf->stream>>>(a));                \
  }
  if (move == MH_RESIMULATE) { if (predrawn) GO(MH_RESIMULATE, true) else GO(MH_RESIMULATE, false) }
  else { if (predrawn) GO(MH_DRIFT, true) else GO(MH_DRIFT, false) }
#undef GO
  CK(cudaGetLastError());
  return GB_OK;
}
template <typename R, bool INIT>
static int launch_mh_k(gb_pf *pf, const MhArgs &a, int move, bool predrawn) {
  switch (pf->model.kind) {
    case MODEL_LGSSM: return launch_mh_t<R, MODEL_LGSSM, INIT>(pf, a, move, predrawn);
    case MODEL_SV: return launch_mh_t<R, MODEL_SV, INIT>(pf, a, move, predrawn);
    case MODEL_BEARINGS: return launch_mh_t<R, MODEL_BEARINGS, INIT>(pf, a, move, predrawn);
    case MODEL_HMM: return launch_mh_t<R, MODEL_HMM, INIT>(pf, a, move, predrawn);
  }
  return fail(GB_EUNSUPPORTED, "gb_pf_rejuvenate: this model kind has no MH rejuvenation moves");
}
// GB_MODEL_SPEC: resimulation move through the NVRTC-specialised Transition (mh_resim_lik_kernel)
static int spec_rejuvenate(gb_pf *pf, int move, const double *obs, int nobs, int64_t *n_accepted, double *alpha_out) {
  const gb_model &m = pf->model;
  if (move != MH_RESIMULATE)
    return fail(GB_EUNSUPPORTED, "gb_pf_rejuvenate: GB_MODEL_SPEC models offer the resimulation move only");
  REQ(nobs == m.nobs && obs, "gb_pf_rejuvenate: wrong number of observations (pass the latest step's observations)");
  if (!pf->mh_ok || pf->pending_gather)
    return fail(GB_ESTATE, "gb_pf_rejuvenate: the parents of the current traces are no longer resident -- call it after "
                           "initialize / particle_filter_step and before maybe_resample (docs/src/tutorials/smc.md:459-466)");
  const bool init = pf->step == 0;
  const int b = init ? 0 : m.n_init_ops, e = init ? m.n_init_ops : m.n_init_ops + m.n_step_ops;
  for (int k = b; k < e; k++)
    if (m.ops[k].code == OP_WEIGHT || m.ops[k].code == OP_LOGPDF_NORMAL || m.ops[k].code == OP_LOGPDF_DIST)
      return fail(GB_EUNSUPPORTED, "gb_pf_rejuvenate: the program carries a custom proposal (WEIGHT / LOGPDF ops); its weight is not "
                                   "the score of the observed choices, which the resimulation move needs");
  JitModule *jm = spec_jit_get(m, pf->dtype);
  const bool predrawn = pf->nn_set > 0 || pf->nu_set > 0;
  const int v = (init ? 2 : 0) + (predrawn ? 1 : 0);
  if (!jm || !jm->fast || !jm->mh[v])
    return fail(GB_EUNSUPPORTED, "gb_pf_rejuvenate: needs the NVRTC-specialised kernels of this program (libnvrtc unavailable, "
                                 "GENB200_SPEC_JIT=0, or a program with gamma draws / more than 16 observations)");
  if (predrawn) REQ(pf->nn_set >= m.nn[init ? 1 : 0] && pf->nu_set >= m.nu[init ? 1 : 0] + 1,
                    "gb_pf_set_noise: a move consumes the program's normal rows and its uniform rows plus ONE accept row");
  if (!pf->lik) CK(gb_dev_alloc(&pf->lik, (size_t)pf->ld * pf->esz));
  if (pf->lik_step != pf->step) {   // scores of the current traces = what the last generate / step added to the weights
    if (!init && pf->incr_step != pf->step) return fail(GB_ESTATE, "gb_pf_rejuvenate: the step's log incremental weights are gone");
    const void *isrc = pf->logw;
    if (!init) { const int irc = incr_source(pf, &isrc); if (irc) return irc; }
    CK(cudaMemcpyAsync(pf->lik, isrc, (size_t)pf->n * pf->esz, cudaMemcpyDeviceToDevice, pf->stream));
    pf->lik_step = pf->step;
  }
  MhArgs a;
  memset(&a, 0, sizeof(a));
  a.mp.obs0 = obs[0];
  for (int k = 1; k < nobs && k < 16; k++) a.mp.obsx[k - 1] = obs[k];
  a.mp.step = (unsigned)pf->step;
  unsigned long long *d_acc = nullptr;
  double *d_alpha = nullptr;
  CK(gb_dev_alloc((void **)&d_acc, 8));
  CK(cudaMemsetAsync(d_acc, 0, 8, pf->stream));
  if (alpha_out) {
    cudaError_t e2 = gb_dev_alloc((void **)&d_alpha, (size_t)pf->n * 8);
    if (e2 != cudaSuccess) { gb_dev_free(d_acc); return fail(GB_ECUDA, cudaGetErrorString(e2)); }
  }
  a.parent = pf->state[pf->cur ^ 1];
  a.cur = pf->state[pf->cur];
  a.anc = (!init && pf->mh_parent_anc) ? pf->anc : nullptr;
  a.normals = pf->d_normals; a.uniforms = pf->d_uniforms;
  a.alpha_out = d_alpha;
  a.lik = pf->lik;
  a.n = pf->n; a.ld = pf->ld; a.pid_offset = pf->pid_offset;
  a.seed = pf->seed; a.step = (unsigned)pf->step; a.move = pf->mh_moves;
  a.accepted = d_acc;
  long long blocks = (pf->n + kBlock - 1) / kBlock;
  const int grid = (int)std::max(1ll, std::min(blocks, (long long)pf->num_sms * jm->mh_occ[v]));
  void *params[] = {&a};
  CUresult cr = CUDA_SUCCESS;
  LAUNCH(pf, GB_K_MISC, cr = g_jit.LaunchKernel(jm->mh[v], grid, 1, 1, kBlock, 1, 1, 0, (CUstream)pf->stream, params, nullptr));
  unsigned long long acc = 0;
  cudaError_t ce = cudaSuccess;
  if (cr == CUDA_SUCCESS) {
    ce = cudaMemcpyAsync(&acc, d_acc, 8, cudaMemcpyDeviceToHost, pf->stream);
    if (ce == cudaSuccess && alpha_out) ce = cudaMemcpyAsync(alpha_out, d_alpha, (size_t)pf->n * 8, cudaMemcpyDeviceToHost, pf->stream);
    if (ce == cudaSuccess && pf->history && !pf->hist_state.empty())
      ce = cudaMemcpy2DAsync(pf->hist_state.back(), (size_t)pf->n_pad * pf->esz, pf->state[pf->cur], (size_t)pf->ld * pf->esz,
                             (size_t)pf->n_pad * pf->esz, pf->S, cudaMemcpyDeviceToDevice, pf->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(pf->stream);
  }
  gb_dev_free(d_acc); gb_dev_free(d_alpha);
  if (cr != CUDA_SUCCESS) {
    const char *es = nullptr;
    g_jit.GetErrorString(cr, &es);
    return fail(GB_ECUDA, std::string("cuLaunchKernel(mh_resim_lik_kernel): ") + (es ? es : "?"));
  }
  if (ce != cudaSuccess) return fail(GB_ECUDA, std::string("gb_pf_rejuvenate: ") + cudaGetErrorString(ce));
  pf->mh_moves += 1;
  pf->nn_set = pf->nu_set = 0;
  if (n_accepted) *n_accepted = (int64_t)acc;
  return GB_OK;
}
extern "C" int gb_pf_rejuvenate(gb_pf_t *pf, int move, const double *obs, int nobs, double drift_std, int64_t *n_accepted,
                                double *alpha_out) {
  REQ(pf, "null handle");
  REQ(move == MH_RESIMULATE || move == MH_DRIFT, "gb_pf_rejuvenate: unknown move");
  const int kind = pf->model.kind;
  if (kind == MODEL_SPEC) return spec_rejuvenate(pf, move, obs, nobs, n_accepted, alpha_out);
  REQ(kind == MODEL_LGSSM || kind == MODEL_SV || kind == MODEL_BEARINGS || kind == MODEL_HMM,
      "gb_pf_rejuvenate: this model kind has no MH rejuvenation moves");
  if (move == MH_DRIFT) {
    if (kind == MODEL_HMM) return fail(GB_EUNSUPPORTED, "gb_pf_rejuvenate: a Gaussian drift move needs continuous latent choices");
    REQ(drift_std > 0.0, "gb_pf_rejuvenate: drift_std must be positive");
  }
  REQ(nobs == pf->model.nobs, "gb_pf_rejuvenate: wrong number of observations (pass the latest step's observations)");
  if (!pf->mh_ok || pf->pending_gather)
    return fail(GB_ESTATE, "gb_pf_rejuvenate: the parents of the current traces are no longer resident -- call it after "
                           "initialize / particle_filter_step and before maybe_resample (docs/src/tutorials/smc.md:459-466)");
  const bool init = pf->step == 0;
  const int nn = move == MH_DRIFT ? (kind == MODEL_BEARINGS ? (init ? 4 : 2) : 1) : model_num_normals(kind, pf->model.params.data(), init);
  const int nu = (move == MH_DRIFT ? 0 : model_num_uniforms(kind)) + 1;
  const bool predrawn = pf->nn_set > 0 || pf->nu_set > 0;
  if (predrawn) REQ(pf->nn_set >= nn && pf->nu_set >= nu, "gb_pf_set_noise: a move consumes the transition's normal rows (drift: one per "
                                                          "latent choice) and its uniform rows plus ONE accept row");
  MhArgs a;
  memset(&a, 0, sizeof(a));
  int rc = fill_model_params(pf, a.mp, obs, nobs, PROPOSAL_DEFAULT, nullptr, 0, init);
  if (rc) return rc;
  unsigned long long *d_acc = nullptr;
  double *d_alpha = nullptr;
  CK(gb_dev_alloc((void **)&d_acc, 8));
  CK(cudaMemsetAsync(d_acc, 0, 8, pf->stream));
  if (alpha_out) {
    cudaError_t e = gb_dev_alloc((void **)&d_alpha, (size_t)pf->n * 8);
    if (e != cudaSuccess) { gb_dev_free(d_acc); return fail(GB_ECUDA, cudaGetErrorString(e)); }
  }
  a.parent = pf->state[pf->cur ^ 1];
  a.cur = pf->state[pf->cur];
  a.anc = (!init && pf->mh_parent_anc) ? pf->anc : nullptr;
  a.normals = pf->d_normals; a.uniforms = pf->d_uniforms;
  a.alpha_out = d_alpha;
  a.n = pf->n; a.ld = pf->ld; a.pid_offset = pf->pid_offset;
  a.seed = pf->seed; a.step = (unsigned)pf->step; a.move = pf->mh_moves;
  a.drift_sd = drift_std;
  a.accepted = d_acc;
  if (init) rc = pf->dtype == GB_F64 ? launch_mh_k<double, true>(pf, a, move, predrawn) : launch_mh_k<float, true>(pf, a, move, predrawn);
  else rc = pf->dtype == GB_F64 ? launch_mh_k<double, false>(pf, a, move, predrawn) : launch_mh_k<float, false>(pf, a, move, predrawn);
  unsigned long long acc = 0;
  cudaError_t e = cudaSuccess;
  if (!rc) {
    e = cudaMemcpyAsync(&acc, d_acc, 8, cudaMemcpyDeviceToHost, pf->stream);
    if (e == cudaSuccess && alpha_out) e = cudaMemcpyAsync(alpha_out, d_alpha, (size_t)pf->n * 8, cudaMemcpyDeviceToHost, pf->stream);
    // the latest history snapshot follows the rejuvenated traces
    if (e == cudaSuccess && pf->history && !pf->hist_state.empty())
      e = cudaMemcpy2DAsync(pf->hist_state.back(), (size_t)pf->n_pad * pf->esz, pf->state[pf->cur], (size_t)pf->ld * pf->esz,
                            (size_t)pf->n_pad * pf->esz, pf->S, cudaMemcpyDeviceToDevice, pf->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(pf->stream);
  }
  gb_dev_free(d_acc); gb_dev_free(d_alpha);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(GB_ECUDA, std::string("gb_pf_rejuvenate: ") + cudaGetErrorString(e));
  pf->mh_moves += 1;
  pf->nn_set = pf->nu_set = 0;
  if (n_accepted) *n_accepted = (int64_t)acc;
  return GB_OK;
}

// Sliding-window rejuvenation (kernels_mh.cuh window_mh_kernel): one block MH move per particle on the latent choices of
// steps a..b, 1 <= a <= b <= current step, reaching back at most W steps (gb_pf_enable_window).
extern "C" int gb_pf_rejuvenate_window(gb_pf_t *pf, int a, int b, const double *obs_window, int nobs, double drift_std,
                                       int64_t *n_accepted, double *alpha_out) {
  REQ(pf && obs_window, "gb_pf_rejuvenate_window: null argument");
  const int kind = pf->model.kind;
  if (!(kind == MODEL_LGSSM || kind == MODEL_SV || kind == MODEL_BEARINGS))
    return fail(GB_EUNSUPPORTED, "gb_pf_rejuvenate_window: Gaussian-drift window moves need continuous latent choices (LGSSM, SV, BEARINGS)");
  REQ(pf->win_W > 0, "gb_pf_rejuvenate_window: call gb_pf_enable_window before initialize");
  REQ(nobs == 1, "gb_pf_rejuvenate_window: one scalar observation per step");
  REQ(drift_std > 0.0, "gb_pf_rejuvenate_window: drift_std must be positive");
  REQ(a >= 1 && a <= b && b <= pf->step, "gb_pf_rejuvenate_window: need 1 <= a <= b <= current step");
  const int L = pf->step - a + 1;
  if (L > pf->win_len)
    return fail(GB_ESTATE, "gb_pf_rejuvenate_window: step a - 1 lies outside the window kept per particle (W = " + std::to_string(pf->win_W) +
                               ", valid lags " + std::to_string(pf->win_len) + ")");
  if (pf->pending_gather) return fail(GB_ESTATE, "gb_pf_rejuvenate_window: call it before maybe_resample (docs/src/tutorials/smc.md:518-524)");
  DeviceScope ds__(pf->device);
  const int nb = b - a + 1;
  const int nd = kind == MODEL_BEARINGS ? 2 : 1;
  const bool predrawn = pf->nn_set > 0 || pf->nu_set > 0;
  if (predrawn) REQ(pf->nn_set >= nb * nd && pf->nu_set >= 1, "gb_pf_set_noise: a window move consumes (b - a + 1) x (choices per step) normal rows and ONE uniform row");
  WinMhArgs wa;
  memset(&wa, 0, sizeof(wa));
  int rc = fill_model_params(pf, wa.mp, obs_window, 1, PROPOSAL_DEFAULT, nullptr, 0, false);
  if (rc) return rc;
  for (int j = 0; j < L; j++) wa.obsw[j] = obs_window[j];
  unsigned long long *d_acc = nullptr;
  double *d_alpha = nullptr;
  CK(gb_dev_alloc((void **)&d_acc, 8));
  CK(cudaMemsetAsync(d_acc, 0, 8, pf->stream));
  if (alpha_out) {
    cudaError_t e = gb_dev_alloc((void **)&d_alpha, (size_t)pf->n * 8);
    if (e != cudaSuccess) { gb_dev_free(d_acc); return fail(GB_ECUDA, cudaGetErrorString(e)); }
  }
  wa.cur = pf->state[pf->cur]; wa.win = pf->win[pf->win_cur];
  wa.normals = pf->d_normals; wa.uniforms = pf->d_uniforms; wa.alpha_out = d_alpha;
  wa.n = pf->n; wa.ld = pf->ld; wa.pid_offset = pf->pid_offset; wa.seed = pf->seed;
  wa.step = (unsigned)pf->step; wa.move = pf->mh_moves;
  wa.L = L; wa.nb = nb; wa.W = pf->win_W; wa.drift_sd = drift_std; wa.accepted = d_acc;
  const int grid = (int)std::min<long long>((pf->n + 127) / 128, (long long)pf->num_sms * 8);
#define GO(R, K)                                                                                              \
  {                                                                                                           \
    if (predrawn) LAUNCH(pf, GB_K_MISC, window_mh_kernel<R, K, true><<<grid, 128, 0, pf->stream>>>(wa));      \
    else LAUNCH(pf, GB_K_MISC, window_mh_kernel<R, K, false><<<grid, 128, 0, pf->stream>>>(wa));              \
  }
  const bool f64 = pf->dtype == GB_F64;
  if (kind == MODEL_LGSSM) { if (f64) GO(double, MODEL_LGSSM) else GO(float, MODEL_LGSSM) }
  else if (kind == MODEL_SV) { if (f64) GO(double, MODEL_SV) else GO(float, MODEL_SV) }
  else { if (f64) GO(double, MODEL_BEARINGS) else GO(float, MODEL_BEARINGS) }
#undef GO
  unsigned long long acc = 0;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(&acc, d_acc, 8, cudaMemcpyDeviceToHost, pf->stream);
  if (e == cudaSuccess && alpha_out) e = cudaMemcpyAsync(alpha_out, d_alpha, (size_t)pf->n * 8, cudaMemcpyDeviceToHost, pf->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(pf->stream);
  gb_dev_free(d_acc); gb_dev_free(d_alpha);
  if (e != cudaSuccess) return fail(GB_ECUDA, std::string("gb_pf_rejuvenate_window: ") + cudaGetErrorString(e));
  pf->mh_moves += 1;
  pf->mh_ok = false;             // the latest-step moves keep their parents in the other trace buffer, which this move bypassed
  pf->nn_set = pf->nu_set = 0;
  if (n_accepted) *n_accepted = (int64_t)acc;
  return GB_OK;
}

// ------------------------------------------------------------------------------------------
// weights: stats, materialisation
// ------------------------------------------------------------------------------------------
static int run_gather(gb_pf *pf);
static int materialize_pending(gb_pf *pf) { return run_gather(pf); }
static int unshare_incr(gb_pf *pf);
static int run_gather(gb_pf *pf) {   // materialise a deferred resample gather
  if (!pf->pending_gather) return GB_OK;
  { const int urc = unshare_incr(pf); if (urc) return urc; }     // (the gather kernel resets logw to 0)
  if (pf->xc_on && pf->a_world > 1) {   // sharded: the peers' offspring rows must have landed before they are read
    LAUNCH(pf, GB_K_MISC, klaunch(xchg_wait_kernel, 1, kMaxWorld, pf->stream, pf->d_stats));
    CK(cudaGetLastError());
  }
  const long long nvec = pf->dtype == GB_F64 ? (pf->n + 1) / 2 : (pf->n + 3) / 4;
  if (pf->dtype == GB_F64) {
    auto kern = gather_kernel<double>;
    int grid = grid_for(pf, kern, nvec, kBlock);
    LAUNCH(pf, GB_K_GATHER, kern<<<grid, kBlock, 0, pf->stream>>>((const double *)pf->state[pf->cur], (double *)pf->state[pf->cur ^ 1],
                                                                   pf->ld, pf->S, pf->anc, pf->n, (double *)pf->logw));
  } else {
    auto kern = gather_kernel<float>;
    int grid = grid_for(pf, kern, nvec, kBlock);
    LAUNCH(pf, GB_K_GATHER, kern<<<grid, kBlock, 0, pf->stream>>>((const float *)pf->state[pf->cur], (float *)pf->state[pf->cur ^ 1],
                                                                   pf->ld, pf->S, pf->anc, pf->n, (float *)pf->logw));
  }
  CK(cudaGetLastError());
  {
    const int wrc = window_update(pf, nullptr, pf->anc, 0);
    if (wrc) return wrc;
  }
  pf->cur ^= 1;                       // particle_filter.jl:270-272
  pf->mh_ok = false;
  pf->pending_gather = false;
  pf->tiles_valid = false;
  return GB_OK;
}

// d_stats->m = max of the log weights (device side)
static int ensure_max(gb_pf *pf, const void *logw) {
  if (pf->wstate != W_DIRTY) return GB_OK;
  if (pf->dtype == GB_F64) {
    auto kern = max_kernel<double>;
    int grid = grid_for(pf, kern, pf->n, kBlock);
    LAUNCH(pf, GB_K_FINALIZE, klaunch(kern, grid, kBlock, pf->stream, (const double *)logw, pf->n, (double *)pf->partials, pf->ticket, pf->d_stats));
  } else {
    auto kern = max_kernel<float>;
    int grid = grid_for(pf, kern, pf->n, kBlock);
    LAUNCH(pf, GB_K_FINALIZE, klaunch(kern, grid, kBlock, pf->stream, (const float *)logw, pf->n, (double *)pf->partials, pf->ticket, pf->d_stats));
  }
  CK(cudaGetLastError());
  pf->wstate = W_MAX;
  pf->max_posted = pf->xc_on;
  return GB_OK;
}
// one pass over the log weights: (s1, s2) w.r.t. the reference max + quantised tile / super-tile totals;
// the reference max is read from d_stats->m (from_device) or passed by value (sharded: global max)
static int run_wstats(gb_pf *pf, const void *logw, double m_arg, bool from_device, bool sync_host = true, RunCtl *ctl = nullptr,
                      const FusedPlan *fused = nullptr, bool exchange = false) {
  FusedPlan fp;
  memset(&fp, 0, sizeof(fp));
  if (fused) fp = *fused;
  const int bits = quant_bits(pf->n_global);
  if (!pf->qbuf) CK(gb_dev_alloc((void **)&pf->qbuf, (size_t)pf->ntiles * kTile * 8));
  if (!pf->ws_clean) {       // the systematic sweep leaves both accumulators zeroed; other paths do not
    CK(cudaMemsetAsync(pf->super_q, 0, (size_t)(pf->nsuper + 1) * 8, pf->stream));
    CK(cudaMemsetAsync(pf->tile_q, 0, (size_t)(pf->ntiles + 1) * 8, pf->stream));
  }
  pf->ws_clean = false;
  // large filters: a warp owns whole tiles (plain stores of the tile totals); otherwise 256/512-particle batches + atomics
  const bool whole = pf->ntiles >= (long long)pf->num_sms * 8 * 4;
  const int mode = exchange ? 2 : (from_device ? 1 : 0);
#define GO(R, W)                                                                                                               \
  {                                                                                                                            \
    auto kern = wstats_kernel<R, W>;                                                                                           \
    const long long units = pf->ntiles * (W ? 1 : kTile / (32 * VecOf<R>::V * 4));                                             \
    int grid = grid_for(pf, kern, units * 32, kBlock);                                                                         \
    LAUNCH(pf, GB_K_FINALIZE, klaunch(kern, grid, kBlock, pf->stream, (const R *)logw, pf->n, m_arg, mode, bits,                \
                                      (unsigned long long *)pf->tile_q, pf->super_q, pf->qbuf, pf->partials, pf->ticket, pf->d_stats, ctl, fp)); \
  }
  if (pf->dtype == GB_F64) { if (whole) GO(double, true) else GO(double, false) }
  else { if (whole) GO(float, true) else GO(float, false) }
#undef GO
  CK(cudaGetLastError());
  if (sync_host) {
    CK(cudaMemcpyAsync(pf->h_stats, pf->d_stats, kStatsRecordBytes, cudaMemcpyDeviceToHost, pf->stream));
    CK(cudaStreamSynchronize(pf->stream));
  } else if (pf->wstate != W_ZERO) {
    pf->wstate = W_MAX;        // the host copy no longer describes the weights; the device max does
  }
  pf->tiles_valid = true;
  return GB_OK;
}
// make h_stats describe the current log weights
static int ensure_stats(gb_pf *pf) {
  if (pf->wstate == W_STATS) return GB_OK;
  if (pf->wstate == W_ZERO) {
    // all log weights are exactly 0.0: logsumexp = 0 + log(N_local)
    pf->h_stats->m = 0.0; pf->h_stats->s1 = (double)pf->n; pf->h_stats->s2 = (double)pf->n;
    pf->h_stats->lse = log((double)pf->n); pf->h_stats->ess = (double)pf->n;
    return GB_OK;
  }
  int rc = ensure_max(pf, pf->logw);
  if (rc) return rc;
  if ((rc = run_wstats(pf, pf->logw, 0.0, true))) return rc;
  pf->wstate = W_STATS;
  return GB_OK;
}

extern "C" int gb_pf_weight_max(gb_pf_t *pf, double *m_local) {
  REQ(pf && m_local, "null argument");
  if (pf->wstate == W_ZERO) { *m_local = 0.0; return GB_OK; }
  if (pf->wstate == W_STATS) { *m_local = pf->h_stats->m; return GB_OK; }
  int rc = ensure_max(pf, pf->logw);
  if (rc) return rc;
  CK(cudaMemcpyAsync(&pf->h_stats->m, &pf->d_stats->m, sizeof(double), cudaMemcpyDeviceToHost, pf->stream));
  CK(cudaStreamSynchronize(pf->stream));
  *m_local = pf->h_stats->m;
  return GB_OK;
}
// Device-resident flavour for the sharded flow (no host round trip between the collectives): the host
// all-reduces stats->m in place, then this launches the statistics pass reading m from device memory.
extern "C" int gb_pf_stats_device(gb_pf_t *pf, void **dev_stats, int *n_words) {
  REQ(pf && dev_stats, "null argument");
  *dev_stats = pf->d_stats;
  if (n_words) *n_words = kStatsRecordBytes / 8;
  return GB_OK;
}
extern "C" int gb_pf_weight_max_device(gb_pf_t *pf) {
  REQ(pf, "null handle");
  if (pf->wstate == W_ZERO) {   // all log weights are exactly 0.0: the device copy of the max must say so
    CK(cudaMemsetAsync(&pf->d_stats->m, 0, sizeof(double), pf->stream));
    return GB_OK;
  }
  if (pf->wstate == W_STATS) {
    CK(cudaMemcpyAsync(&pf->d_stats->m, &pf->h_stats->m, sizeof(double), cudaMemcpyHostToDevice, pf->stream));
    return GB_OK;
  }
  return ensure_max(pf, pf->logw);
}
extern "C" int gb_pf_weight_sums_device(gb_pf_t *pf, int bits) {
  REQ(pf, "null handle");
  REQ(bits == quant_bits(pf->n_global), "gb_pf_weight_sums_device: bits must be gb_quant_bits(n_global)");
  int rc = materialize_pending(pf);
  if (rc) return rc;
  return run_wstats(pf, pf->logw, 0.0, true, false);
}
extern "C" int gb_pf_weight_sums(gb_pf_t *pf, double m_ref, int bits, double *s1, double *s2, uint64_t *q_local) {
  REQ(pf && s1 && s2 && q_local, "null argument");
  REQ(bits == quant_bits(pf->n_global), "gb_pf_weight_sums: bits must be gb_quant_bits(n_global)");
  int rc = materialize_pending(pf);
  if (rc) return rc;
  if ((rc = run_wstats(pf, pf->logw, m_ref, false))) return rc;
  if (pf->wstate != W_ZERO) pf->wstate = W_STATS;    // h_stats now refers to m_ref (>= the local max)
  *s1 = pf->h_stats->s1; *s2 = pf->h_stats->s2; *q_local = pf->h_stats->q_total;
  return GB_OK;
}
extern "C" int gb_pf_get_ess(gb_pf_t *pf, double *ess, double *lse) {
  REQ(pf, "null handle");
  int rc = ensure_stats(pf);
  if (rc) return rc;
  if (ess) *ess = pf->h_stats->ess;
  if (lse) *lse = pf->h_stats->lse;
  return GB_OK;
}
extern "C" int gb_pf_log_ml_estimate(gb_pf_t *pf, double *out) {
  REQ(pf && out, "null argument");
  REQ(pf->n == pf->n_global, "gb_pf_log_ml_estimate: sharded filters combine partials on the host (sharded.py)");
  int rc = ensure_stats(pf);
  if (rc) return rc;
  // particle_filter.jl:93: log_ml_est + logsumexp(log_weights) - log(num_particles)
  *out = pf->log_ml_est + pf->h_stats->lse - log((double)pf->n_global);
  return GB_OK;
}
extern "C" int gb_pf_get_log_ml_est(gb_pf_t *pf, double *out) { REQ(pf && out, "null argument"); *out = pf->log_ml_est; return GB_OK; }

template <typename R> static int fill_dev(gb_pf *pf, R *p, long long n, R v) {
  auto kern = fill_kernel<R>;
  int grid = grid_for(pf, kern, n, kBlock);
  LAUNCH(pf, GB_K_MISC, kern<<<grid, kBlock, 0, pf->stream>>>(p, n, v));
  CK(cudaGetLastError());
  return GB_OK;
}
// deferred state -> the arrays really hold what the API says they hold
static int materialize(gb_pf *pf) { return run_gather(pf); /* also zeroes logw */ }

// Device -> caller memory for the N-element results (log weights, trace columns, parents, normalised IS weights).
// The caller's buffer is pageable and usually untouched (a fresh Julia / numpy array): a plain cudaMemcpy then runs at
// 3-4 GB/s (80 MB of log weights: 21 ms, three times the kernel that produced them).  Large copies are therefore staged:
// DMA into two cached pinned chunks at link speed while a few host threads move the previous chunk into the caller's
// pages (and widen float -> double for fp32 filters on the way).
template <typename Src, typename Dst>
static void host_fan_out(Dst *dst, const Src *src, size_t n) {
  const unsigned hw = std::thread::hardware_concurrency();
  const int T = (int)std::max(1u, std::min(8u, hw ? hw : 1u));
  auto work = [=](int t) {
    const size_t lo = n * t / T, hi = n * (t + 1) / T;
    if (std::is_same<Src, Dst>::value) memcpy((void *)(dst + lo), (const void *)(src + lo), (hi - lo) * sizeof(Dst));
    else for (size_t i = lo; i < hi; i++) dst[i] = (Dst)src[i];
  };
  if (T == 1 || n < (1u << 18)) { for (int t = 0; t < T; t++) work(t); return; }
  std::vector<std::thread> th;
  for (int t = 1; t < T; t++) th.emplace_back(work, t);
  work(0);
  for (auto &x : th) x.join();
}
template <typename Src, typename Dst>
static int staged_d2h(cudaStream_t stream, const Src *dsrc, Dst *out, size_t n) {
  constexpr size_t kChunkBytes = 16u << 20;
  const size_t per = kChunkBytes / sizeof(Src);
  if (n * sizeof(Src) < (4u << 20)) {       // small: one plain copy (+ widening)
    if (std::is_same<Src, Dst>::value) {
      CK(cudaMemcpyAsync(out, dsrc, n * sizeof(Src), cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
    } else {
      std::vector<Src> tmp(n);
      CK(cudaMemcpyAsync(tmp.data(), dsrc, n * sizeof(Src), cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      for (size_t i = 0; i < n; i++) out[i] = (Dst)tmp[i];
    }
    return GB_OK;
  }
  Src *pin[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaError_t e = gb_host_alloc(&pin[0], kChunkBytes);
  if (e == cudaSuccess) e = gb_host_alloc(&pin[1], kChunkBytes);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
  const size_t nchunks = (n + per - 1) / per;
  for (size_t k = 0; k <= nchunks && e == cudaSuccess; k++) {
    if (k < nchunks) {
      const size_t off = k * per, cnt = std::min(per, n - off);
      e = cudaMemcpyAsync(pin[k & 1], dsrc + off, cnt * sizeof(Src), cudaMemcpyDeviceToHost, stream);
      if (e == cudaSuccess) e = cudaEventRecord(ev[k & 1], stream);
    }
    if (k > 0 && e == cudaSuccess) {        // chunk k is in flight while chunk k-1 lands in the caller's pages
      const size_t off = (k - 1) * per, cnt = std::min(per, n - off);
      e = cudaEventSynchronize(ev[(k - 1) & 1]);
      if (e == cudaSuccess) host_fan_out(out + off, pin[(k - 1) & 1], cnt);
    }
  }
  if (e != cudaSuccess) cudaStreamSynchronize(stream);
  {
    FreeBatch fb__;
    gb_host_free(pin[0]); gb_host_free(pin[1]);
  }
  if (ev[0]) cudaEventDestroy(ev[0]);
  if (ev[1]) cudaEventDestroy(ev[1]);
  if (e != cudaSuccess) return fail(GB_ECUDA, std::string("device -> host copy: ") + cudaGetErrorString(e));
  return GB_OK;
}
static int copy_out(gb_pf *pf, const void *dsrc, double *out, long long n) {
  if (pf->dtype == GB_F64) return staged_d2h(pf->stream, (const double *)dsrc, out, (size_t)n);
  return staged_d2h(pf->stream, (const float *)dsrc, out, (size_t)n);
}
extern "C" int gb_pf_get_log_weights(gb_pf_t *pf, double *out) {
  REQ(pf && out, "null argument");
  int rc = materialize(pf);
  if (rc) return rc;
  return copy_out(pf, pf->logw, out, pf->n);
}
// the log incremental weights of the LAST particle_filter_step! (its return value, particle_filter.jl:199-207, :231-239),
// kept on the device until the next step so that a caller who never looks at them does not pay the N-element copy
extern "C" int gb_pf_get_log_increments(gb_pf_t *pf, double *out) {
  REQ(pf && out, "null argument");
  if (pf->step < 1 || pf->incr_step != pf->step)
    return fail(GB_ESTATE, "gb_pf_get_log_increments: no particle_filter_step has produced increments for the current step");
  const void *src = nullptr;
  int rc = incr_source(pf, &src);
  if (rc) return rc;
  return copy_out(pf, src, out, pf->n);
}
extern "C" int gb_pf_set_log_weights(gb_pf_t *pf, const double *in) {
  REQ(pf && in, "null argument");
  int rc = materialize(pf);
  if (rc) return rc;
  if ((rc = unshare_incr(pf))) return rc;
  if (pf->dtype == GB_F64) {
    CK(cudaMemcpyAsync(pf->logw, in, (size_t)pf->n * 8, cudaMemcpyHostToDevice, pf->stream));
  } else {
    std::vector<float> tmp(pf->n);
    for (long long i = 0; i < pf->n; i++) tmp[i] = (float)in[i];
    CK(cudaMemcpyAsync(pf->logw, tmp.data(), (size_t)pf->n * 4, cudaMemcpyHostToDevice, pf->stream));
  }
  CK(cudaStreamSynchronize(pf->stream));
  pf->wstate = W_DIRTY;
  pf->max_posted = false;
  pf->tiles_valid = false;
  return GB_OK;
}
extern "C" int gb_pf_get_state(gb_pf_t *pf, int field, double *out) {
  REQ(pf && out, "null argument");
  REQ(field >= 0 && field < pf->S, "gb_pf_get_state: field out of range");
  int rc = materialize(pf);
  if (rc) return rc;
  return copy_out(pf, (const char *)pf->state[pf->cur] + (size_t)field * pf->ld * pf->esz, out, pf->n);
}
// parents64 = Gen's `parents` (global, 1-based) of the current traces
static int ensure_parents64(gb_pf *pf) {
  if (!pf->parents64) CK(gb_dev_alloc((void **)&pf->parents64, (size_t)pf->ld * 8));
  if (pf->parents64_valid) return GB_OK;
  if (pf->anc_valid) {
    if (pf->xc_on && pf->a_world > 1) {   // sharded: ancestors / parent ids of imported slots arrive by peer stores (phase C)
      LAUNCH(pf, GB_K_MISC, klaunch(xchg_wait_kernel, 1, kMaxWorld, pf->stream, pf->d_stats));
      CK(cudaGetLastError());
    }
    int grid = grid_for(pf, parents_from_anc_kernel, pf->n, kBlock);
    LAUNCH(pf, GB_K_MISC, parents_from_anc_kernel<<<grid, kBlock, 0, pf->stream>>>(pf->anc, pf->parents64, pf->n, pf->pid_offset,
                                                                                       pf->ext_used ? pf->n_pad : (long long)1 << 40, pf->ext_parents));
  } else {
    int grid = grid_for(pf, iota_parents_kernel, pf->n, kBlock);
    LAUNCH(pf, GB_K_MISC, iota_parents_kernel<<<grid, kBlock, 0, pf->stream>>>(pf->parents64, pf->n, pf->pid_offset));
  }
  CK(cudaGetLastError());
  pf->parents64_valid = true;
  return GB_OK;
}
extern "C" int gb_pf_get_parents(gb_pf_t *pf, int64_t *out) {
  REQ(pf && out, "null argument");
  int rc = ensure_parents64(pf);
  if (rc) return rc;
  return staged_d2h(pf->stream, (const long long *)pf->parents64, (long long *)out, (size_t)pf->n);
}

extern "C" int gb_pf_clone(gb_pf_t *pf, gb_pf_t **out) {
  REQ(pf && out, "null argument");
  int rc = materialize(pf);
  if (rc) return rc;
  if ((rc = unshare_incr(pf))) return rc;
  rc = ensure_stats(pf);
  if (rc) return rc;
  if ((rc = ensure_parents64(pf))) return rc;
  gb_pf_t *c = nullptr;
  rc = gb_pf_create(&pf->model, pf->n, pf->n_global, pf->pid_offset, pf->dtype, pf->seed, &c);
  if (rc) return rc;
  c->stream = pf->stream;
  c->lazy = pf->lazy;
  c->step = pf->step;
  c->log_ml_est = pf->log_ml_est;
  c->wstate = pf->wstate;
  *c->h_stats = *pf->h_stats;      // W_STATS: the host record describes the (copied) weights
  c->incr_step = pf->incr_step;
  c->anc_valid = false;            // parents travel as the materialised int64 array
  size_t col = (size_t)pf->ld * pf->esz;
  cudaError_t e = cudaMemcpyAsync(c->state[c->cur], pf->state[pf->cur], col * pf->S, cudaMemcpyDeviceToDevice, pf->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(c->logw, pf->logw, col, cudaMemcpyDeviceToDevice, pf->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(c->incr, pf->incr, col, cudaMemcpyDeviceToDevice, pf->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(c->anc, pf->anc, (size_t)pf->ld * 4, cudaMemcpyDeviceToDevice, pf->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_stats, pf->d_stats, kStatsRecordBytes, cudaMemcpyDeviceToDevice, pf->stream);
  if (e == cudaSuccess && pf->parents64_valid) {
    e = gb_dev_alloc((void **)&c->parents64, (size_t)pf->ld * 8);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->parents64, pf->parents64, (size_t)pf->ld * 8, cudaMemcpyDeviceToDevice, pf->stream);
    c->parents64_valid = true;
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(pf->stream);
  if (e != cudaSuccess) { pf_release(c); return fail(GB_ECUDA, std::string("gb_pf_clone: ") + cudaGetErrorString(e)); }
  *out = c;
  return GB_OK;
}

// ------------------------------------------------------------------------------------------
// resampling
// ------------------------------------------------------------------------------------------
// inclusive integer CDF of the current weights into pf->cdf (needs valid tile totals: run_wstats); tile_q becomes its prefix
static int launch_cdf(gb_pf *pf) {
  if (!pf->cdf) CK(gb_dev_alloc((void **)&pf->cdf, (size_t)pf->ld * 8));
  LAUNCH(pf, GB_K_TILESCAN, tilescan_cdf_kernel<<<1, 1024, 0, pf->stream>>>(pf->ntiles, pf->tile_q, pf->d_plan, pf->n_global));
  CK(cudaGetLastError());
  LAUNCH(pf, GB_K_ANCESTORS, cdf_kernel<<<(unsigned)pf->ntiles, kBlock, 0, pf->stream>>>(pf->qbuf, pf->tile_q, pf->cdf));
  CK(cudaGetLastError());
  pf->tiles_valid = false;
  pf->ws_clean = false;
  return GB_OK;
}

// systematic pipeline after wstats: superscan -> ancestors -> overflow.  Fills anc[k - anc_base] for the
// shard's offspring slots; single shard: q_offset = 0, q_total = 0 (= the local total), anc_base = 0.
template <typename R>
static int launch_systematic(gb_pf *pf, const void *logw, double m, uint64_t q_offset, uint64_t q_total, uint32_t u0,
                             const AncOut &ao, const RunCtl *ctl = nullptr, int plan_preset = 0, bool plan_done = false,
                             bool export_rows = false) {
  const int bits = quant_bits(pf->n_global);
  if (!plan_done) {   // (fused run: the statistics kernel's last CTA already scanned the super-tiles and wrote the plan)
    LAUNCH(pf, GB_K_TILESCAN, klaunch(superscan_kernel, 1, 1024, pf->stream, pf->nsuper, pf->super_q, pf->super_excl, pf->d_plan, q_offset, q_total,
                                      pf->n_global, u0, pf->oflow_count, plan_preset));
  }
  CK(cudaGetLastError());
  LAUNCH(pf, GB_K_ANCESTORS, klaunch(ancestors_sys_kernel, (unsigned)pf->ntiles, kBlock, pf->stream, (const unsigned long long *)pf->qbuf,
                                     (const uint64_t *)pf->tile_q, (const unsigned long long *)pf->super_excl, (const ResamplePlan *)pf->d_plan, ao,
                                     pf->oflow, pf->oflow_count, ctl));
  CK(cudaGetLastError());
  ExportArgs ex;
  memset(&ex, 0, sizeof(ex));
  if (export_rows) {       // sharded sweep: offspring owned by other shards leave in the same kernel (peer stores over NVLink)
    ex.sin = pf->state[pf->cur]; ex.ld = pf->ld; ex.S = pf->S; ex.cur = pf->cur; ex.spill = pf->off_anc; ex.sh = pf->d_shplan;
    ex.offs = pf->d_offs; ex.pid_offset = pf->pid_offset; ex.spill_cap = pf->off_cap; ex.stats = pf->d_stats;
    ex.grid_barrier = pf->oflow_count + 8;
  }
  {
    auto okern = overflow_sys_kernel<R>;
    const unsigned ogrid = (unsigned)grid_for(pf, okern, (long long)pf->num_sms * 4 * kBlock, kBlock);   // co-resident by construction
    LAUNCH(pf, export_rows ? GB_K_GATHER : GB_K_ANCESTORS,
           klaunch(okern, ogrid, kBlock, pf->stream, (const unsigned long long *)pf->qbuf, (const ResamplePlan *)pf->d_plan, ao,
                   (const OverflowEntry *)pf->oflow, (const unsigned int *)pf->oflow_count, (unsigned long long *)pf->tile_q, pf->ntiles, ctl, ex));
  }
  CK(cudaGetLastError());
  pf->tiles_valid = false;   // superscan / overflow zeroed the accumulators for the next statistics pass
  pf->ws_clean = true;
  return GB_OK;
}

// ancestors (local, 0-based int32) of all n slots of a single-shard filter into anc.
// Requires h_stats (m) and, for systematic, valid tile totals (run_wstats).
template <typename R>
static int compute_ancestors(gb_pf *pf, const void *logw, double m, int kind, uint32_t u0, const uint64_t *d_u64,
                             uint64_t seed, uint32_t step, int *anc) {
  int rc;
  if (kind == GB_RESAMPLE_SYSTEMATIC || kind == GB_RESAMPLE_RESIDUAL) {
    // residual resampling (deterministic copies + a systematic sweep over the residual masses, same u0) IS systematic
    // resampling in the exact integer spec: kernels_resample.cuh "RESIDUAL == SYSTEMATIC"
    if (!pf->tiles_valid && (rc = run_wstats(pf, logw, m, false))) return rc;
    return launch_systematic<R>(pf, logw, m, 0, 0, u0, AncOut{anc, 0, pf->n_global, nullptr, 0, 0});
  }
  if (kind == GB_RESAMPLE_MULTINOMIAL) {
    if (!pf->tiles_valid && (rc = run_wstats(pf, logw, m, false))) return rc;
    // guide table first: the systematic sweep for u0 = 0 consumes (and zeroes) the tile totals, so they are parked meanwhile
    if (!pf->guide) CK(gb_dev_alloc((void **)&pf->guide, (size_t)pf->ld * 4));
    if (!pf->tile_save) CK(gb_dev_alloc((void **)&pf->tile_save, (size_t)(pf->ntiles + 1) * 8));
    CK(cudaMemcpyAsync(pf->tile_save, pf->tile_q, (size_t)(pf->ntiles + 1) * 8, cudaMemcpyDeviceToDevice, pf->stream));
    if ((rc = launch_systematic<R>(pf, logw, m, 0, 0, 0u, AncOut{pf->guide, 0, pf->n_global, nullptr, 0, 0}))) return rc;
    CK(cudaMemcpyAsync(pf->tile_q, pf->tile_save, (size_t)(pf->ntiles + 1) * 8, cudaMemcpyDeviceToDevice, pf->stream));
    if ((rc = launch_cdf(pf))) return rc;
    int grid = grid_for(pf, multinomial_guided_kernel, pf->n, kBlock);
    LAUNCH(pf, GB_K_ANCESTORS, multinomial_guided_kernel<<<grid, kBlock, 0, pf->stream>>>(pf->cdf, pf->guide, pf->n, pf->d_plan, d_u64, seed,
                                                                                          step, pf->n, anc, (unsigned)PK_MULTI));
    CK(cudaGetLastError());
    return GB_OK;
  }
  if (kind == GB_RESAMPLE_MULTINOMIAL_SORTED) {
    // sorted multinomial: exponential spacings -> sorted thresholds -> one merge with the integer CDF (kernels_msort.cuh).
    // d_u64: n + 1 caller-supplied spacings on the device (raw-array entry), or ms_e already holds them (ms_e_set)
    if (!pf->tiles_valid && (rc = run_wstats(pf, logw, m, false))) return rc;
    if ((rc = msort_workspace(pf))) return rc;
    const long long nst = msort_slot_tiles(pf), nslot = (pf->n + kTile - 1) / kTile;
    unsigned long long *st_sum = pf->ms_st, *st_pref = pf->ms_st + nst + 1;
    if (d_u64) CK(cudaMemcpyAsync(pf->ms_e, d_u64, (size_t)(pf->n + 1) * 8, cudaMemcpyDeviceToDevice, pf->stream));
    const bool given = d_u64 || pf->ms_e_set;
    if (pf->ms_e_async) {            // spacings + slot-tile sums were started on the auxiliary stream: join
      CK(cudaStreamWaitEvent(pf->stream, pf->ms_join, 0));
      pf->ms_e_async = false;
    } else {
      const int grid = (int)std::min<long long>(nst, (long long)pf->num_sms * 8);
      if (given)
        LAUNCH(pf, GB_K_MISC, msort_spacing_kernel<false><<<grid, kBlock, 0, pf->stream>>>(pf->ms_e, pf->n, nst, 0ull, 0u, 0ll, st_sum));
      else
        LAUNCH(pf, GB_K_MISC, msort_spacing_kernel<true><<<grid, kBlock, 0, pf->stream>>>(pf->ms_e, pf->n, nst, seed, step, 0ll, st_sum));
    }
    LAUNCH(pf, GB_K_TILESCAN, msort_scan_kernel<<<2, 1024, 0, pf->stream>>>(st_sum, st_pref, nst, (const unsigned long long *)pf->tile_q, pf->ms_tcdf,
                                                                             pf->ntiles, pf->ms_plan));
    {
      const int grid = (int)std::min<long long>((nst + 1 + kBlock - 1) / kBlock, (long long)pf->num_sms * 8);
      LAUNCH(pf, GB_K_TILESCAN, msort_partition_kernel<<<grid, kBlock, 0, pf->stream>>>(st_pref, nst, pf->ms_tcdf, pf->ntiles, pf->ms_plan, pf->ms_pt));
    }
    {
      const size_t smem = (size_t)kMsPad * sizeof(uint64_t);          // the CDF of kMsTiles particle tiles: 36.9 KB for two
      static bool attr_set[64] = {false};
      const int dv = pf->device >= 0 && pf->device < 64 ? pf->device : 0;
      if (!attr_set[dv]) {
        CK(cudaFuncSetAttribute(msort_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dv] = true;
      }
      LAUNCH(pf, GB_K_ANCESTORS, msort_fill_kernel<<<(unsigned)nslot, kBlock, smem, pf->stream>>>(pf->ms_e, st_pref, (const unsigned long long *)pf->qbuf,
                                                                                                 pf->ms_tcdf, pf->ntiles, pf->ms_pt, pf->ms_plan, pf->n, anc));
    }
    CK(cudaGetLastError());
    pf->tiles_valid = false;      // (the tile totals were only read: the next statistics pass clears them itself)
    pf->ws_clean = false;
    pf->ms_e_set = false;
    return GB_OK;
  }
  return fail(GB_EINVAL, "unknown resampler kind");
}

extern "C" int gb_pf_maybe_resample(gb_pf_t *pf, double ess_threshold, int kind, int *did) {
  REQ(pf && did, "null argument");
  REQ(pf->n == pf->n_global, "gb_pf_maybe_resample: single shard only; sharded filters use the gb_pf_* pieces");
  *did = 0;
  int rc = ensure_stats(pf);          // particle_filter.jl:254-255
  if (rc) return rc;
  const double ess = pf->h_stats->ess, lse = pf->h_stats->lse, m = pf->h_stats->m;
  const bool do_resample = ess < ess_threshold;   // :256 (strict; NaN compares false)
  if (!do_resample) return GB_OK;
  if ((rc = materialize_pending(pf))) return rc;  // two deferred gathers cannot be chained
  // (W_ZERO with nothing pending: the array was really zeroed by the gather / commit that set it)
  // resampling randomness is keyed by the index of the step that produced the weights
  const uint32_t u0 = pf->use_u0 ? pf->u0 : philox_block(pf->seed, 0, (uint32_t)pf->step, PK_RESAMPLE, 0).x;
  const uint64_t *d_u64 = pf->u64_set ? pf->d_u64 : nullptr;
  rc = pf->dtype == GB_F64 ? compute_ancestors<double>(pf, pf->logw, m, kind, u0, d_u64, pf->seed, (uint32_t)pf->step, pf->anc)
                           : compute_ancestors<float>(pf, pf->logw, m, kind, u0, d_u64, pf->seed, (uint32_t)pf->step, pf->anc);
  if (rc) return rc;
  pf->log_ml_est += lse - log((double)pf->n_global);   // :263
  pf->hist_resampled = true;
  pf->mh_ok = false;              // `anc` now describes the NEXT generation
  pf->anc_valid = true;
  pf->parents64_valid = false;
  pf->pending_gather = true;
  pf->wstate = W_ZERO;                                  // :266
  pf->tiles_valid = false;
  pf->use_u0 = false;
  pf->u64_set = false;
  if (!pf->lazy && (rc = run_gather(pf))) return rc;    // :264-272
  *did = 1;
  return GB_OK;
}

// sample_unweighted_traces (particle_filter.jl:101-109): n_samples iid indices from Categorical(normalised weights)
// (integer CDF + Philox, stream PK_SAMPLE keyed by (seed, draw)), and the corresponding trace rows
extern "C" int gb_pf_sample_unweighted(gb_pf_t *pf, int64_t n_samples, uint64_t draw, int64_t *idx_out, double *rows_out) {
  REQ(pf && n_samples >= 1 && idx_out, "gb_pf_sample_unweighted: bad argument");
  REQ(pf->n == pf->n_global, "gb_pf_sample_unweighted: single shard only");
  int rc = materialize_pending(pf);
  if (rc) return rc;
  if ((rc = ensure_stats(pf))) return rc;
  const double m = pf->h_stats->m;
  if (!(m > -INFINITY)) return fail(GB_ESTATE, "gb_pf_sample_unweighted: all log weights are -Inf or NaN");
  if (!pf->tiles_valid && (rc = run_wstats(pf, pf->logw, m, false))) return rc;
  if ((rc = launch_cdf(pf))) return rc;
  int *d_idx = nullptr;
  long long *d_idx1 = nullptr;
  double *d_rows = nullptr;
  cudaError_t e = gb_dev_alloc((void **)&d_idx, (size_t)n_samples * 4);
  if (e == cudaSuccess) e = gb_dev_alloc((void **)&d_idx1, (size_t)n_samples * 8);
  if (e == cudaSuccess && rows_out) e = gb_dev_alloc((void **)&d_rows, (size_t)n_samples * pf->S * 8);
  if (e == cudaSuccess) {
    int grid = grid_for(pf, multinomial_kernel, n_samples, kBlock);
    LAUNCH(pf, GB_K_MISC, multinomial_kernel<<<grid, kBlock, 0, pf->stream>>>(pf->cdf, pf->tile_q, pf->ntiles, pf->n, pf->d_plan, nullptr,
                                                                              pf->seed, (unsigned)draw, n_samples, d_idx, (unsigned)PK_SAMPLE));
    if (pf->dtype == GB_F64)
      LAUNCH(pf, GB_K_MISC, sample_rows_kernel<double><<<grid, kBlock, 0, pf->stream>>>((const double *)pf->state[pf->cur], pf->ld, rows_out ? pf->S : 0,
                                                                                        d_idx, n_samples, d_rows, d_idx1, pf->pid_offset));
    else
      LAUNCH(pf, GB_K_MISC, sample_rows_kernel<float><<<grid, kBlock, 0, pf->stream>>>((const float *)pf->state[pf->cur], pf->ld, rows_out ? pf->S : 0,
                                                                                       d_idx, n_samples, d_rows, d_idx1, pf->pid_offset));
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(idx_out, d_idx1, (size_t)n_samples * 8, cudaMemcpyDeviceToHost, pf->stream);
  if (e == cudaSuccess && rows_out) e = cudaMemcpyAsync(rows_out, d_rows, (size_t)n_samples * pf->S * 8, cudaMemcpyDeviceToHost, pf->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(pf->stream);
  gb_dev_free(d_idx); gb_dev_free(d_idx1); gb_dev_free(d_rows);
  if (e != cudaSuccess) return fail(GB_ECUDA, std::string("gb_pf_sample_unweighted: ") + cudaGetErrorString(e));
  return GB_OK;
}

// ---- fused multi-step entry: T x { maybe_resample!(systematic); particle_filter_step! } with the ESS test, the
//      log-ML update and the gathered/direct choice made on the device (no host round trip per step) ----
// ---- small particle counts: the whole T-step loop in one single-CTA kernel (kernels_small.cuh) ----
static bool small_run_eligible(const gb_pf *pf, int T) {
  static const bool on = [] { const char *e = getenv("GENB200_SMALL_RUN"); return !(e && e[0] == '0'); }();
  const int k = pf->model.kind;
  return on && T > 0 && pf->n <= kSmallMax && pf->n == pf->n_global && pf->pid_offset == 0 && !pf->history &&
         (k == MODEL_LGSSM || k == MODEL_SV || k == MODEL_BEARINGS || k == MODEL_HMM);
}
static int small_run(gb_pf *pf, int T, const double *obs, int nobs, int *n_resampled) {
  SmallRunArgs a;
  memset(&a, 0, sizeof(a));
  int rc = GB_OK;
  for (int t = 0; t < T && !rc; t++)   // per-step argument checks (HMM: observation in 1..M), parameters from the last one
    rc = fill_model_params(pf, a.mp, obs + (size_t)t * nobs, nobs, PROPOSAL_DEFAULT, nullptr, 0, false);
  if (rc) return rc;
  double *d_obs = nullptr;
  CK(gb_dev_alloc((void **)&d_obs, (size_t)T * nobs * 8));
  cudaError_t e = cudaMemcpyAsync(d_obs, obs, (size_t)T * nobs * 8, cudaMemcpyHostToDevice, pf->stream);
  a.obs = d_obs; a.T = T; a.nobs = nobs;
  a.state[0] = pf->state[0]; a.state[1] = pf->state[1]; a.cur = pf->cur;
  a.logw = pf->logw; a.incr = pf->incr; a.anc = pf->anc;
  a.n = pf->n; a.ld = pf->ld; a.seed = pf->seed; a.step = (unsigned)pf->step;
  a.bits = quant_bits(pf->n_global);
  a.ctl = pf->d_ctl; a.stats = pf->d_stats;
  if (e == cudaSuccess) {
#define GO(R, K) LAUNCH(pf, GB_K_STEP, small_run_kernel<R, K><<<1, kSmallThreads, 0, pf->stream>>>(a))
    const bool f64 = pf->dtype == GB_F64;
    switch (pf->model.kind) {
      case MODEL_LGSSM: if (f64) GO(double, MODEL_LGSSM); else GO(float, MODEL_LGSSM); break;
      case MODEL_SV: if (f64) GO(double, MODEL_SV); else GO(float, MODEL_SV); break;
      case MODEL_BEARINGS: if (f64) GO(double, MODEL_BEARINGS); else GO(float, MODEL_BEARINGS); break;
      default: if (f64) GO(double, MODEL_HMM); else GO(float, MODEL_HMM); break;
    }
#undef GO
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(pf->h_ctl, pf->d_ctl, sizeof(RunCtl), cudaMemcpyDeviceToHost, pf->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(pf->stream);
  gb_dev_free(d_obs);
  if (e != cudaSuccess) return fail(GB_ECUDA, std::string("gb_pf_run (single-CTA path): ") + cudaGetErrorString(e));
  pf->log_ml_est = pf->h_ctl->log_ml_est;
  pf->step += T;
  pf->cur ^= (T & 1);
  pf->incr_step = pf->step;
  pf->incr_where = 0;
  if (pf->h_ctl->n_resampled > 0) { pf->anc_valid = true; pf->parents64_valid = false; pf->ext_used = false; }
  pf->wstate = W_MAX; pf->tiles_valid = false; pf->pending_gather = false;
  if (n_resampled) *n_resampled = pf->h_ctl->n_resampled;
  return GB_OK;
}

// ---- mid-size particle counts: the whole T-step loop in one persistent co-resident kernel (kernels_mid.cuh) ----
static long long mid_run_max_particles() {
  const char *e = getenv("GENB200_MID_RUN_MAX");      // (read per call: the tests pin either path)
  return e ? atoll(e) : 2000000ll;     // beyond this the HBM-streaming kernels win: N = 2e6 55 vs 59 us/step, 4e6 101 vs 102 (fp32 88 vs 84)
}
static bool mid_run_eligible(const gb_pf *pf, int T) {
  const int k = pf->model.kind;
  return T > 0 && pf->n > kSmallMax && pf->n <= mid_run_max_particles() && pf->n == pf->n_global && pf->pid_offset == 0 && !pf->history &&
         (k == MODEL_LGSSM || k == MODEL_SV || k == MODEL_BEARINGS || k == MODEL_HMM);
}
static int mid_run(gb_pf *pf, int T, const double *obs, int nobs, int *n_resampled) {
  MidRunArgs a;
  memset(&a, 0, sizeof(a));
  int rc = GB_OK;
  for (int t = 0; t < T && !rc; t++)   // per-step argument checks (HMM: observation in 1..M), parameters from the last one
    rc = fill_model_params(pf, a.mp, obs + (size_t)t * nobs, nobs, PROPOSAL_DEFAULT, nullptr, 0, false);
  if (rc) return rc;
  const void *kern = nullptr;
  const bool f64 = pf->dtype == GB_F64;
#define PICK(K, B) kern = f64 ? (const void *)mid_run_kernel<double, K, B> : (const void *)mid_run_kernel<float, K, B>
  switch (pf->model.kind) {
    case MODEL_LGSSM: PICK(MODEL_LGSSM, 4); break;
    case MODEL_SV: PICK(MODEL_SV, 4); break;
    case MODEL_BEARINGS: PICK(MODEL_BEARINGS, 3); break;
    default: PICK(MODEL_HMM, 4); break;
  }
#undef PICK
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBlock, 0));
  REQ(occ >= 1, "gb_pf_run: the persistent kernel does not fit on an SM");
  // co-resident grid, every SM equally full: the step phase deals the particles evenly over the CTAs, whatever the tile count
  // (small filters: no more CTAs than eight per tile, a CTA needs a few hundred particles to be worth a barrier seat)
  int grid = pf->num_sms * occ;
  while (grid > pf->num_sms && (long long)grid > pf->ntiles * 8) grid -= pf->num_sms;
  if (!pf->mid_ws || pf->mid_grid < grid) {
    gb_dev_free(pf->mid_ws);
    pf->mid_ws = nullptr;
    CK(gb_dev_alloc(&pf->mid_ws, 256 + (size_t)grid * 8 + 256 + (size_t)grid * sizeof(MidPartial)));
    pf->mid_grid = grid;
  }
  const size_t bar_bytes = (256 + (size_t)grid * 8 + 255) / 256 * 256;      // release word (own line) + arrival words
  if (!pf->qbuf) CK(gb_dev_alloc((void **)&pf->qbuf, (size_t)pf->ntiles * kTile * 8));
  double *d_obs = nullptr;
  CK(gb_dev_alloc((void **)&d_obs, (size_t)T * nobs * 8));
  cudaError_t e = cudaMemcpyAsync(d_obs, obs, (size_t)T * nobs * 8, cudaMemcpyHostToDevice, pf->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(pf->mid_ws, 0, bar_bytes, pf->stream);
  a.obs = d_obs; a.T = T; a.nobs = nobs;
  a.state[0] = pf->state[0]; a.state[1] = pf->state[1]; a.cur = pf->cur;
  a.logw = pf->logw; a.incr = pf->incr; a.anc = pf->anc;
  a.qbuf = pf->qbuf;
  a.barrier.release = (unsigned long long *)pf->mid_ws;
  a.barrier.arrive = (unsigned long long *)((char *)pf->mid_ws + 256);
  a.part = (MidPartial *)((char *)pf->mid_ws + bar_bytes);
  a.n = pf->n; a.ld = pf->ld; a.ntiles = pf->ntiles; a.seed = pf->seed; a.step = (unsigned)pf->step;
  a.bits = quant_bits(pf->n_global);
  a.ctl = pf->d_ctl; a.stats = pf->d_stats;
  static const bool dbg = [] { const char *v = getenv("GENB200_MID_DEBUG"); return v && v[0] == '1'; }();
  unsigned long long *d_dbg = nullptr;
  if (dbg && e == cudaSuccess) {
    e = gb_dev_alloc((void **)&d_dbg, (size_t)T * 64);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_dbg, 0, (size_t)T * 64, pf->stream);
    a.dbg = d_dbg;
  }
  if (e == cudaSuccess) {
    void *params[] = {(void *)&a};
    LAUNCH(pf, GB_K_STEP, e = cudaLaunchCooperativeKernel(kern, dim3((unsigned)grid), dim3(kBlock), params, 0, pf->stream));
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(pf->h_ctl, pf->d_ctl, sizeof(RunCtl), cudaMemcpyDeviceToHost, pf->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(pf->stream);
  if (d_dbg && e == cudaSuccess) {      // phase breakdown of CTA 0 (ns per step, averaged over the run)
    std::vector<unsigned long long> h((size_t)T * 8);
    e = cudaMemcpy(h.data(), d_dbg, (size_t)T * 64, cudaMemcpyDeviceToHost);
    double ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = 0; t < T; t++) {
      for (int k = 0; k < 6; k++) ph[k] += (double)(h[(size_t)t * 8 + k + 1] - h[(size_t)t * 8 + k]);
      ph[7] += (double)(h[(size_t)t * 8 + 7] - h[(size_t)t * 8 + 6]);
      if (t + 1 < T) ph[6] += (double)(h[(size_t)(t + 1) * 8] - h[(size_t)t * 8 + 7]);
    }
    fprintf(stderr, "[genb200 persistent run] grid %d x %d, n %lld, T %d; ns/step: stats %.0f | barrier %.0f | ancestors %.0f | barrier %.0f | step %.0f | "
                    "barrier %.0f | (bare barrier %.0f) | max fold %.0f\n", grid, kBlock, pf->n, T, ph[0] / T, ph[1] / T, ph[2] / T, ph[3] / T, ph[4] / T,
            ph[5] / T, ph[7] / T, ph[6] / std::max(T - 1, 1));
  }
  gb_dev_free(d_dbg);
  gb_dev_free(d_obs);
  if (e != cudaSuccess) return fail(GB_ECUDA, std::string("gb_pf_run (persistent path): ") + cudaGetErrorString(e));
  pf->log_ml_est = pf->h_ctl->log_ml_est;
  pf->step += T;
  pf->cur ^= (T & 1);
  pf->incr_step = pf->step;
  pf->incr_where = 0;
  if (pf->h_ctl->n_resampled > 0) { pf->anc_valid = true; pf->parents64_valid = false; pf->ext_used = false; }
  pf->wstate = W_MAX; pf->tiles_valid = false; pf->pending_gather = false;
  if (n_resampled) *n_resampled = pf->h_ctl->n_resampled;
  return GB_OK;
}

// T x { maybe_resample!; particle_filter_step!; [sweeps x mh on the latest latents] } without a host round trip
static int pf_run_impl(gb_pf *pf, int T, const double *obs, int nobs, double ess_threshold, int move, double drift_std, int sweeps,
                       int *n_resampled, int64_t *n_accepted) {
  if (pf) pf->mh_ok = false;
  REQ(pf && obs && T >= 0, "gb_pf_run: bad argument");
  REQ(pf->n == pf->n_global, "gb_pf_run: single shard only");
  REQ(pf->model.kind != MODEL_BLR, "the BLR model has no sequential structure (importance sampling only)");
  REQ(nobs == pf->model.nobs, "gb_pf_run: wrong number of observations per step");
  REQ(pf->nn_set == 0 && pf->nu_set == 0 && !pf->use_u0, "gb_pf_run: validation noise is not supported by the fused entry");
  int rc = materialize_pending(pf);
  if (rc) return rc;
  if ((rc = gb_pf_weight_max_device(pf))) return rc;     // d_stats->m current for the first statistics pass
  if (!pf->d_ctl) {
    CK(gb_dev_alloc((void **)&pf->d_ctl, sizeof(RunCtl)));
    CK(gb_host_alloc((void **)&pf->h_ctl, sizeof(RunCtl)));
  }
  pf->h_ctl->do_resample = 0; pf->h_ctl->n_resampled = 0;
  pf->h_ctl->log_ml_est = pf->log_ml_est;
  pf->h_ctl->ess_threshold = ess_threshold;
  pf->h_ctl->log_n = log((double)pf->n_global);
  CK(cudaMemcpyAsync(pf->d_ctl, pf->h_ctl, sizeof(RunCtl), cudaMemcpyHostToDevice, pf->stream));
  CK(cudaStreamSynchronize(pf->stream));                // h_ctl is reused for the read-back below
  pf->win_len = 0;                    // (the fused entries do not maintain the rejuvenation window)
  if (sweeps == 0 && small_run_eligible(pf, T)) return small_run(pf, T, obs, nobs, n_resampled);
  if (sweeps == 0 && mid_run_eligible(pf, T)) return mid_run(pf, T, obs, nobs, n_resampled);
  if (sweeps > 0) {
    const int kind = pf->model.kind;
    REQ(move == MH_RESIMULATE || move == MH_DRIFT, "gb_pf_run_mh: unknown move");
    if (!(kind == MODEL_LGSSM || kind == MODEL_SV || kind == MODEL_BEARINGS || kind == MODEL_HMM) || (move == MH_DRIFT && kind == MODEL_HMM))
      return fail(GB_EUNSUPPORTED, "gb_pf_run_mh: this model kind does not have this MH rejuvenation move");
    REQ(move != MH_DRIFT || drift_std > 0.0, "gb_pf_run_mh: drift_std must be positive");
  }
  // every step's arguments are checked BEFORE the first launch (as the single-CTA path does): a bad observation at step t > 0
  // must not leave the handle half way through the run
  if (pf->model.kind == MODEL_SPEC) {
    REQ(pf->model.n_step_ops > 0, "this SPEC model has no step program");
  } else {
    ModelParams mp_check;
    for (int t = 0; t < T; t++)
      if ((rc = fill_model_params(pf, mp_check, obs + (size_t)t * nobs, nobs, PROPOSAL_DEFAULT, nullptr, 0, false))) return rc;
  }
  unsigned long long *d_acc = nullptr;
  if (sweeps > 0) {
    CK(gb_dev_alloc((void **)&d_acc, 8));
    cudaError_t me = cudaMemsetAsync(d_acc, 0, 8, pf->stream);
    if (me != cudaSuccess) { gb_dev_free(d_acc); return fail(GB_ECUDA, std::string("gb_pf_run_mh: ") + cudaGetErrorString(me)); }
  }
  const AncOut ao{pf->anc, 0, pf->n_global, nullptr, 0, 0};
  int done = 0;          // steps fully enqueued
  for (int t = 0; t < T && !rc; t++) {
    // maybe_resample!: statistics + decision on the device, then the (conditional) systematic sweep
    const uint32_t u0 = philox_block(pf->seed, 0, (uint32_t)pf->step, PK_RESAMPLE, 0).x;
    const FusedPlan fp{pf->d_plan, pf->super_excl, pf->oflow_count, pf->n_global, u0};
    if ((rc = run_wstats(pf, pf->logw, 0.0, true, false, pf->d_ctl, &fp))) break;
    rc = pf->dtype == GB_F64 ? launch_systematic<double>(pf, pf->logw, 0.0, 0, 0, u0, ao, pf->d_ctl, 0, true)
                             : launch_systematic<float>(pf, pf->logw, 0.0, 0, 0, u0, ao, pf->d_ctl, 0, true);
    if (rc) break;
    // particle_filter_step!
    if ((rc = launch_model(pf, obs + (size_t)t * nobs, nobs, PROPOSAL_DEFAULT, nullptr, 0, false, false, 2))) break;
    pf->step += 1;
    pf->cur ^= 1;
    done = t + 1;
    // rejuvenation of the traces just extended (docs/src/tutorials/smc.md:459-463, :518-520): their parents are the
    // buffer the step read from, through `anc` iff the device decided to resample in this iteration
    for (int sw = 0; sw < sweeps && !rc; sw++) {
      MhArgs a;
      memset(&a, 0, sizeof(a));
      if ((rc = fill_model_params(pf, a.mp, obs + (size_t)t * nobs, nobs, PROPOSAL_DEFAULT, nullptr, 0, false))) break;
      a.parent = pf->state[pf->cur ^ 1];
      a.cur = pf->state[pf->cur];
      a.anc = pf->anc; a.ctl = pf->d_ctl;
      a.n = pf->n; a.ld = pf->ld; a.pid_offset = pf->pid_offset;
      a.seed = pf->seed; a.step = (unsigned)pf->step; a.move = (unsigned)sw;
      a.drift_sd = drift_std;
      a.accepted = d_acc;
      rc = pf->dtype == GB_F64 ? launch_mh_k<double, false>(pf, a, move, false) : launch_mh_k<float, false>(pf, a, move, false);
    }
  }
  // Read the control block back and bring the host-side bookkeeping in line with whatever WAS enqueued -- also when a
  // launch failed part of the way (the device has run `done` complete steps and possibly the resampling half of one more)
  unsigned long long acc = 0;
  cudaError_t ce = cudaMemcpyAsync(pf->h_ctl, pf->d_ctl, sizeof(RunCtl), cudaMemcpyDeviceToHost, pf->stream);
  if (ce == cudaSuccess && d_acc) ce = cudaMemcpyAsync(&acc, d_acc, 8, cudaMemcpyDeviceToHost, pf->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(pf->stream);
  gb_dev_free(d_acc);
  if (ce == cudaSuccess && (done > 0 || rc)) {
    // (a launch that failed after validation is a CUDA error, i.e. fatal for the context; the cached statistics are dropped)
    pf->wstate = rc ? W_DIRTY : W_MAX;
    pf->tiles_valid = false; pf->ws_clean = false; pf->pending_gather = false;
    if (done > 0) pf->incr_step = pf->step;
    if (pf->h_ctl->n_resampled > 0) { pf->anc_valid = true; pf->parents64_valid = false; pf->ext_used = false; }
    pf->log_ml_est = pf->h_ctl->log_ml_est;
  }
  if (rc) return rc;
  if (ce != cudaSuccess) return fail(GB_ECUDA, std::string("gb_pf_run: ") + cudaGetErrorString(ce));
  if (n_accepted) *n_accepted = (int64_t)acc;
  if (n_resampled) *n_resampled = pf->h_ctl->n_resampled;
  return GB_OK;
}
extern "C" int gb_pf_run(gb_pf_t *pf, int T, const double *obs, int nobs, double ess_threshold, int *n_resampled) {
  return pf_run_impl(pf, T, obs, nobs, ess_threshold, 0, 0.0, 0, n_resampled, nullptr);
}
extern "C" int gb_pf_run_mh(gb_pf_t *pf, int T, const double *obs, int nobs, double ess_threshold, int move, double drift_std, int sweeps,
                            int *n_resampled, int64_t *n_accepted) {
  REQ(sweeps >= 0 && sweeps <= 1024, "gb_pf_run_mh: sweeps must be in 0..1024");
  return pf_run_impl(pf, T, obs, nobs, ess_threshold, move, drift_std, sweeps, n_resampled, n_accepted);
}

// ---- sharded pieces ----
extern "C" int gb_pf_systematic_offspring(gb_pf_t *pf, double m_global, int bits, uint64_t q_offset, uint64_t q_local,
                                          uint64_t q_total, uint32_t u0, int64_t *slot_lo, int64_t *slot_hi) {
  if (pf) pf->mh_ok = false;
  REQ(pf && slot_lo && slot_hi, "null argument");
  REQ(q_total > 0, "gb_pf_systematic_offspring: zero total weight");
  REQ(bits == quant_bits(pf->n_global), "gb_pf_systematic_offspring: bits must be gb_quant_bits(n_global)");
  REQ(pf->tiles_valid, "gb_pf_systematic_offspring: call gb_pf_weight_sums(m_global) first");
  int rc;
  // the shard's offspring range is [slots_below(q_offset), slots_below(q_offset + Q_local)): host-side
  // evaluation of the same exact integer function the kernels use
  SweepParams sp;
  sp.q_total = q_total; sp.q_inv = 1.0 / (double)q_total; sp.dscale = (uint64_t)pf->n_global << 32;
  sp.ratio = (double)pf->n_global / (double)q_total; sp.u0 = u0;
  const long long lo = slots_below(q_offset, sp), hi = slots_below(q_offset + q_local, sp);
  const long long cnt = hi - lo;
  if (cnt > pf->off_cap) {
    gb_dev_free(pf->off_anc);
    pf->off_anc = nullptr;
    // grow with head-room: the offspring count fluctuates from resample to resample and every
    // cudaFree/cudaMalloc is a device synchronisation
    long long want = cnt + cnt / 4 + kTile;
    if (want < pf->n_pad + pf->n_pad / 4) want = pf->n_pad + pf->n_pad / 4;
    long long cap = ((want + kTile - 1) / kTile) * kTile;
    CK(gb_dev_alloc((void **)&pf->off_anc, (size_t)cap * 4));
    pf->off_cap = cap;
  }
  if (cnt > 0) {
    // offspring slots this shard owns go straight into its `anc` column; the rest into the spill buffer
    const AncOut ao{pf->anc, pf->pid_offset, pf->pid_offset + pf->n, pf->off_anc, lo, pf->off_cap};
    rc = pf->dtype == GB_F64 ? launch_systematic<double>(pf, pf->logw, m_global, q_offset, q_total, u0, ao)
                             : launch_systematic<float>(pf, pf->logw, m_global, q_offset, q_total, u0, ao);
    if (rc) return rc;
  }
  pf->off_lo = lo; pf->off_hi = hi;
  pf->ext_used = false;
  *slot_lo = lo; *slot_hi = hi;
  return GB_OK;
}
// ---- asynchronous sharded run: the exchange plan, the resample decision and the log-ML live on the device; offspring
//      owned by other shards are written straight into their memory (CUDA IPC peer pointers over NVLink); the host only
//      enqueues (the collectives in between are issued by gen.jl_b200/sharded.py on the same stream) ----
struct IpcBundle { cudaIpcMemHandle_t h[5]; long long ld, n_pad, ext_cap; };
static int ensure_async_alloc(gb_pf *pf) {
  if (!pf->d_ctl) {
    CK(gb_dev_alloc((void **)&pf->d_ctl, sizeof(RunCtl)));
    CK(gb_host_alloc((void **)&pf->h_ctl, sizeof(RunCtl)));
  }
  if (!pf->d_shplan) {
    CK(gb_dev_alloc((void **)&pf->d_shplan, sizeof(ShardPlanDev)));
    CK(gb_host_alloc((void **)&pf->h_shplan, sizeof(ShardPlanDev)));
    memset(pf->h_shplan, 0, sizeof(ShardPlanDev));
    CK(gb_dev_alloc((void **)&pf->d_offs, sizeof(long long) * (kMaxWorld + 1)));
  }
  if (!pf->mail) {
    CK(gb_dev_alloc((void **)&pf->mail, sizeof(XchgMail)));
    CK(cudaMemset(pf->mail, 0, sizeof(XchgMail)));
    CK(gb_dev_alloc((void **)&pf->d_xc, sizeof(XchgCtl)));
    CK(gb_host_alloc((void **)&pf->h_xc, sizeof(XchgCtl)));
    memset(pf->h_xc, 0, sizeof(XchgCtl));
  }
  return GB_OK;
}
extern "C" int gb_pf_ipc_export(gb_pf_t *pf, void *bundle_out, int *nbytes) {
  REQ(pf && bundle_out && nbytes, "null argument");
  REQ(pf->ext_parents, "gb_pf_ipc_export: not a sharded filter (n_local == n_global)");
  DeviceScope ds__(pf->device);
  int rc = ensure_async_alloc(pf);
  if (rc) return rc;
  IpcBundle b;
  memset(&b, 0, sizeof(b));
  CK(cudaIpcGetMemHandle(&b.h[0], pf->state[0]));
  CK(cudaIpcGetMemHandle(&b.h[1], pf->state[1]));
  CK(cudaIpcGetMemHandle(&b.h[2], pf->anc));
  CK(cudaIpcGetMemHandle(&b.h[3], pf->ext_parents));
  CK(cudaIpcGetMemHandle(&b.h[4], pf->mail));
  b.ld = pf->ld; b.n_pad = pf->n_pad; b.ext_cap = pf->ext_cap;
  memcpy(bundle_out, &b, sizeof(b));
  *nbytes = (int)sizeof(b);
  return GB_OK;
}
extern "C" int gb_pf_ipc_bundle_bytes(void) { return (int)sizeof(IpcBundle); }
// peer in ANOTHER process: open its IPC bundle
extern "C" int gb_pf_async_attach_ipc(gb_pf_t *pf, int peer_rank, const void *bundle) {
  REQ(pf && bundle && peer_rank >= 0 && peer_rank < kMaxWorld, "gb_pf_async_attach_ipc: bad argument");
  DeviceScope ds__(pf->device);
  int rc = ensure_async_alloc(pf);
  if (rc) return rc;
  IpcBundle b;
  memcpy(&b, bundle, sizeof(b));
  PeerMem &pm = pf->h_shplan->peer[peer_rank];
  void *p[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int k = 0; k < 5; k++) {
    CK(cudaIpcOpenMemHandle(&p[k], b.h[k], cudaIpcMemLazyEnablePeerAccess));
    pf->ipc_opened.push_back(p[k]);
  }
  pm.state[0] = p[0]; pm.state[1] = p[1]; pm.anc = (int *)p[2]; pm.ext_parents = (long long *)p[3];
  pm.ld = b.ld; pm.n_pad = b.n_pad; pm.ext_cap = b.ext_cap;
  pf->h_xc->peer[peer_rank] = (XchgMail *)p[4];
  return GB_OK;
}
// peer in THIS process (another device with peer access enabled, or the same device)
extern "C" int gb_pf_async_attach_local(gb_pf_t *pf, int peer_rank, gb_pf_t *peer) {
  REQ(pf && peer && peer_rank >= 0 && peer_rank < kMaxWorld, "gb_pf_async_attach_local: bad argument");
  int rc;
  { DeviceScope ds__(pf->device); if ((rc = ensure_async_alloc(pf))) return rc; }
  { DeviceScope ds__(peer->device); if ((rc = ensure_async_alloc(peer))) return rc; }
  if (peer->device != pf->device) {
    DeviceScope ds__(pf->device);
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, pf->device, peer->device));
    if (!can) return fail(GB_EUNSUPPORTED, "gb_pf_async_attach_local: no peer access between the two devices");
    cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
    cudaGetLastError();
  }
  PeerMem &pm = pf->h_shplan->peer[peer_rank];
  pm.state[0] = peer->state[0]; pm.state[1] = peer->state[1]; pm.anc = peer->anc; pm.ext_parents = peer->ext_parents;
  pm.ld = peer->ld; pm.n_pad = peer->n_pad; pm.ext_cap = peer->ext_cap;
  pf->h_xc->peer[peer_rank] = peer->mail;
  return GB_OK;
}
static long long xchg_timeout_cycles() {
  const char *e = getenv("GENB200_XCHG_TIMEOUT_MS");
  const double ms = e ? atof(e) : 20000.0;
  return (long long)(ms * 1.9e6);       // ~1.9 GHz SM clock
}
// ONCE per filter, after every peer has been attached: shard geometry, the exchange control block and the mailbox.
// Resets this shard's mailbox and post counters, so the caller must put a barrier ACROSS ALL SHARDS (device idle on every
// shard) between this call and the first gb_pf_xchg_stats / gb_pf_async_iterate.
extern "C" int gb_pf_async_setup(gb_pf_t *pf, int world, int rank, const int64_t *offs, double ess_threshold) {
  REQ(pf && offs, "null argument");
  REQ(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "gb_pf_async_setup: world size must be <= 64");
  REQ(offs[rank] == pf->pid_offset && offs[rank + 1] == pf->pid_offset + pf->n && offs[world] == pf->n_global,
      "gb_pf_async_setup: shard offsets do not match this shard");
  DeviceScope ds__(pf->device);
  int rc = materialize_pending(pf);
  if (rc) return rc;
  if ((rc = ensure_async_alloc(pf))) return rc;
  for (int d = 0; d < world; d++)
    if (d != rank) REQ(pf->h_shplan->peer[d].anc != nullptr && pf->h_xc->peer[d] != nullptr,
                       "gb_pf_async_setup: attach every peer first (gb_pf_async_attach_*)");
  // both trace buffers must be in lock step with the peers': cur parity is part of the protocol
  pf->h_shplan->world = world; pf->h_shplan->rank = rank; pf->h_shplan->overflow = 0;
  CK(cudaMemcpyAsync(pf->d_shplan, pf->h_shplan, sizeof(ShardPlanDev), cudaMemcpyHostToDevice, pf->stream));
  std::vector<long long> o(offs, offs + world + 1);
  CK(cudaMemcpyAsync(pf->d_offs, o.data(), sizeof(long long) * (world + 1), cudaMemcpyHostToDevice, pf->stream));
  XchgCtl *xc = pf->h_xc;
  xc->mine = pf->mail; xc->peer[rank] = pf->mail;
  xc->world = world; xc->rank = rank;
  xc->seq_a = xc->seq_b = xc->seq_c = 0;
  xc->timeout_cycles = xchg_timeout_cycles();
  xc->error = 0; xc->export_ticket = 0;
  CK(cudaMemsetAsync(pf->mail, 0, sizeof(XchgMail), pf->stream));
  CK(cudaMemcpyAsync(pf->d_xc, xc, sizeof(XchgCtl), cudaMemcpyHostToDevice, pf->stream));
  CK(cudaMemcpyAsync(&pf->d_stats->xc, &pf->d_xc, sizeof(XchgCtl *), cudaMemcpyHostToDevice, pf->stream));
  pf->xc_on = true;
  pf->max_posted = false;
  // spill buffer: ancestors of the offspring that land outside this shard's own slot range
  const long long want = pf->n_pad + pf->n_pad / 4 + kTile;
  if (pf->off_cap < want) {
    gb_dev_free(pf->off_anc);
    pf->off_anc = nullptr;
    CK(gb_dev_alloc((void **)&pf->off_anc, (size_t)want * 4));
    pf->off_cap = want;
  }
  CK(cudaStreamSynchronize(pf->stream));
  pf->a_world = world; pf->a_rank = rank;
  return gb_pf_async_arm(pf, ess_threshold);
}
// per run / per synchronous maybe_resample!: the device-side control block takes over the host's log_ml_est and threshold
extern "C" int gb_pf_async_arm(gb_pf_t *pf, double ess_threshold) {
  REQ(pf && pf->a_world > 0, "gb_pf_async_arm: call gb_pf_async_setup first");
  DeviceScope ds__(pf->device);
  pf->h_ctl->do_resample = 0; pf->h_ctl->n_resampled = 0;
  pf->h_ctl->log_ml_est = pf->log_ml_est;
  pf->h_ctl->ess_threshold = ess_threshold;
  pf->h_ctl->log_n = log((double)pf->n_global);
  CK(cudaMemcpyAsync(pf->d_ctl, pf->h_ctl, sizeof(RunCtl), cudaMemcpyHostToDevice, pf->stream));
  CK(cudaStreamSynchronize(pf->stream));       // h_ctl is reused for the read-back
  return GB_OK;
}
extern "C" int gb_pf_async_cur(const gb_pf_t *pf, int *cur) { REQ(pf && cur, "null argument"); *cur = pf->cur; return GB_OK; }
// Phases A + B of the exchange for the current weights: (post this shard's max if the kernel that produced the weights did
// not: gb_pf_xchg_post) -> statistics pass w.r.t. the GLOBAL max, whose last CTA posts (sum e, sum e^2, integer CDF mass) to
// every shard.  A driver with several shards on ONE stream calls gb_pf_xchg_post for all of them first.
extern "C" int gb_pf_xchg_post(gb_pf_t *pf) {
  REQ(pf && pf->a_world > 0 && pf->xc_on, "gb_pf_xchg_post: call gb_pf_async_setup first");
  DeviceScope ds__(pf->device);
  int rc = materialize_pending(pf);
  if (rc) return rc;
  if (!pf->max_posted) {
    if (pf->wstate == W_ZERO) {     // all log weights are exactly 0.0
      CK(cudaMemsetAsync(&pf->d_stats->m, 0, sizeof(double), pf->stream));
      LAUNCH(pf, GB_K_MISC, xchg_post_max_kernel<<<1, 32, 0, pf->stream>>>(pf->d_stats));
      CK(cudaGetLastError());
      pf->max_posted = true;
    } else {                        // (d_stats->m may hold a previous GLOBAL max: recompute the shard's own)
      pf->wstate = W_DIRTY;
      if ((rc = ensure_max(pf, pf->logw))) return rc;
    }
  }
  return GB_OK;
}
extern "C" int gb_pf_xchg_stats(gb_pf_t *pf) {
  int rc = gb_pf_xchg_post(pf);
  if (rc) return rc;
  DeviceScope ds__(pf->device);
  const int keep = pf->wstate;
  rc = run_wstats(pf, pf->logw, 0.0, true, false, nullptr, nullptr, true);
  pf->wstate = keep == W_ZERO ? W_ZERO : W_MAX;   // d_stats->m is now the global max: still a valid reference for this shard
  return rc;
}
// gb_pf_xchg_stats + gb_pf_async_plan(u0, armed threshold, do_scan = 1) as ONE kernel: the statistics kernel's last CTA waits
// for the peers' records and computes the plan in its tail.  Only for shards that do not share a device with a peer (the
// waiting CTA would sit in front of the peer's statistics kernel on a shared stream).
extern "C" int gb_pf_xchg_stats_plan(gb_pf_t *pf, uint32_t u0) {
  int rc = gb_pf_xchg_post(pf);
  if (rc) return rc;
  if (pf) pf->mh_ok = false;
  DeviceScope ds__(pf->device);
  const int keep = pf->wstate;
  FusedPlan fp;
  memset(&fp, 0, sizeof(fp));
  fp.plan = pf->d_plan; fp.super_excl = pf->super_excl; fp.overflow_count = pf->oflow_count; fp.n_global = pf->n_global; fp.u0 = u0;
  fp.sh = pf->d_shplan; fp.offs = pf->d_offs; fp.spill_cap = pf->off_cap; fp.sh_ctl = pf->d_ctl;
  rc = run_wstats(pf, pf->logw, 0.0, true, false, nullptr, &fp, true);
  pf->wstate = keep == W_ZERO ? W_ZERO : W_MAX;
  return rc;
}
// ESS test + log-ML update + exchange plan on the device from the G records in this shard's mailbox.  ess_threshold NaN:
// the armed threshold.  do_scan != 0 prepares the sweep that follows (consumes the super-tile totals).
extern "C" int gb_pf_async_plan(gb_pf_t *pf, uint32_t u0, double ess_threshold, int do_scan) {
  if (pf) pf->mh_ok = false;
  REQ(pf && pf->a_world > 0 && pf->xc_on, "gb_pf_async_plan: call gb_pf_async_setup first");
  REQ(!do_scan || pf->tiles_valid, "gb_pf_async_plan: statistics pass missing (gb_pf_xchg_stats)");
  DeviceScope ds__(pf->device);
  LAUNCH(pf, GB_K_TILESCAN, klaunch(shard_plan_kernel, 1, 1024, pf->stream, pf->d_stats, (const long long *)pf->d_offs, pf->n_global, u0, ess_threshold,
                                    do_scan, pf->off_cap, pf->d_ctl, pf->d_plan, pf->d_shplan, pf->nsuper, pf->super_q, pf->super_excl,
                                    pf->oflow_count));
  CK(cudaGetLastError());
  return GB_OK;
}
extern "C" int gb_pf_async_sweep(gb_pf_t *pf) {
  if (pf) pf->mh_ok = false;
  REQ(pf && pf->a_world > 0 && pf->tiles_valid, "gb_pf_async_sweep: statistics pass missing");
  DeviceScope ds__(pf->device);
  const AncOut ao{pf->anc, pf->pid_offset, pf->pid_offset + pf->n, pf->off_anc, -1, pf->off_cap};
  const bool ex = pf->a_world > 1;      // offspring owned by other shards leave with the sweep's second kernel (+ phase C post)
  return pf->dtype == GB_F64 ? launch_systematic<double>(pf, pf->logw, 0.0, 0, 0, 0, ao, pf->d_ctl, 1, true, ex)
                             : launch_systematic<float>(pf, pf->logw, 0.0, 0, 0, 0, ao, pf->d_ctl, 1, true, ex);
}
// (kept for drivers written against the first version of the protocol) the export of the offspring owned by other shards now
// rides in the second kernel of gb_pf_async_sweep (overflow_sys_kernel): one launch less per step
extern "C" int gb_pf_async_export(gb_pf_t *pf) {
  REQ(pf && pf->a_world > 0, "gb_pf_async_export: bad argument");
  return GB_OK;
}
// (kept for drivers written against the first version of the protocol) the wait for the peers' "done" flags now sits at
// the top of the step kernel itself (xchg_wait_done): one launch less per step
extern "C" int gb_pf_async_wait(gb_pf_t *pf) {
  REQ(pf && pf->a_world > 0, "gb_pf_async_wait: bad argument");
  return GB_OK;
}
extern "C" int gb_pf_async_step(gb_pf_t *pf, const double *obs, int nobs) {
  if (pf) pf->mh_ok = false;
  REQ(pf && obs && pf->a_world > 0, "gb_pf_async_step: bad argument");
  REQ(nobs == pf->model.nobs, "gb_pf_async_step: wrong number of observations");
  DeviceScope ds__(pf->device);
  int rc = launch_model(pf, obs, nobs, PROPOSAL_DEFAULT, nullptr, 0, false, false, 2);
  if (rc) return rc;
  pf->step += 1;
  pf->cur ^= 1;
  pf->wstate = W_MAX;
  pf->max_posted = pf->xc_on;
  pf->tiles_valid = false;
  pf->incr_step = pf->step;
  return GB_OK;
}
// one whole iteration { maybe_resample!(armed threshold); particle_filter_step!(obs) } of ONE shard, nothing but enqueues
extern "C" int gb_pf_async_iterate(gb_pf_t *pf, const double *obs, int nobs) {
  REQ(pf && pf->a_world > 0, "gb_pf_async_iterate: call gb_pf_async_setup first");
  const uint32_t u0 = philox_block(pf->seed, 0, (uint32_t)pf->step, PK_RESAMPLE, 0).x;
  int rc = gb_pf_xchg_stats_plan(pf, u0);          // (one shard per process: no peer shares this device)
  if (!rc) rc = gb_pf_async_sweep(pf);
  if (!rc) rc = gb_pf_async_export(pf);
  if (!rc) rc = gb_pf_async_wait(pf);
  if (!rc) rc = gb_pf_async_step(pf, obs, nobs);
  return rc;
}
static int xchg_check(gb_pf *pf) {      // after a host sync: capacity overflow / exchange timeout flags
  if (pf->h_shplan->overflow)
    return fail(GB_ESTATE, "sharded resample: an exchange capacity was exceeded (weights too skewed for the extension rows); "
                           "the filter state is invalid -- use the staged synchronous path");
  if (pf->h_xc->error)
    return fail(GB_ESTATE, "sharded resample: a peer did not answer within GENB200_XCHG_TIMEOUT_MS (shards out of lock step?)");
  return GB_OK;
}
extern "C" int gb_pf_async_finish(gb_pf_t *pf, int *n_resampled, int *overflow) {
  REQ(pf && pf->a_world > 0, "gb_pf_async_finish: bad argument");
  DeviceScope ds__(pf->device);
  CK(cudaMemcpyAsync(pf->h_ctl, pf->d_ctl, sizeof(RunCtl), cudaMemcpyDeviceToHost, pf->stream));
  CK(cudaMemcpyAsync(pf->h_shplan, pf->d_shplan, sizeof(ShardPlanDev), cudaMemcpyDeviceToHost, pf->stream));
  CK(cudaMemcpyAsync(&pf->h_xc->error, &pf->d_xc->error, sizeof(int), cudaMemcpyDeviceToHost, pf->stream));
  CK(cudaStreamSynchronize(pf->stream));
  pf->log_ml_est = pf->h_ctl->log_ml_est;
  if (pf->h_ctl->n_resampled > 0) { pf->anc_valid = true; pf->parents64_valid = false; pf->ext_used = pf->a_world > 1; }
  pf->pending_gather = false;
  if (n_resampled) *n_resampled = pf->h_ctl->n_resampled;
  if (overflow) *overflow = pf->h_shplan->overflow | (pf->h_xc->error << 1);
  return xchg_check(pf);
}
// Global weight statistics {m, sum e, sum e^2, lse, ess, Q} over ALL shards as the last gb_pf_async_plan saw them, and its
// decision; one device -> host read.
extern "C" int gb_pf_xchg_read(gb_pf_t *pf, double *lse, double *ess, uint64_t *q_total, int *do_resample) {
  REQ(pf && pf->a_world > 0, "gb_pf_xchg_read: bad argument");
  DeviceScope ds__(pf->device);
  CK(cudaMemcpyAsync(pf->h_ctl, pf->d_ctl, sizeof(RunCtl), cudaMemcpyDeviceToHost, pf->stream));
  {   // only what the plan kernel writes for the host: {world, rank, overflow, global record, per-shard masses}
    const size_t off = offsetof(ShardPlanDev, world);
    CK(cudaMemcpyAsync((char *)pf->h_shplan + off, (const char *)pf->d_shplan + off, sizeof(ShardPlanDev) - off, cudaMemcpyDeviceToHost, pf->stream));
  }
  CK(cudaMemcpyAsync(&pf->h_xc->error, &pf->d_xc->error, sizeof(int), cudaMemcpyDeviceToHost, pf->stream));
  CK(cudaStreamSynchronize(pf->stream));
  if (lse) *lse = pf->h_shplan->global.lse;
  if (ess) *ess = pf->h_shplan->global.ess;
  if (q_total) *q_total = pf->h_shplan->global.q_total;
  if (do_resample) *do_resample = pf->h_ctl->do_resample;
  return xchg_check(pf);
}
// (no device access) the global max and every shard's integer CDF mass as of the last gb_pf_xchg_read
extern "C" int gb_pf_xchg_records(gb_pf_t *pf, double *m_global, uint64_t *q_shard) {
  REQ(pf && pf->a_world > 0 && pf->h_shplan, "gb_pf_xchg_records: bad argument");
  if (m_global) *m_global = pf->h_shplan->global.m;
  if (q_shard) for (int r = 0; r < pf->a_world; r++) q_shard[r] = pf->h_shplan->q_shard[r];
  return GB_OK;
}
// Synchronous maybe_resample! over the exchange: after plan(threshold) -> sweep -> export -> wait, read the device's
// decision and do the host-side bookkeeping of particle_filter.jl:263-272 (the gather stays deferred).
extern "C" int gb_pf_xchg_commit(gb_pf_t *pf, int *did_resample, double *ess) {
  REQ(pf && did_resample, "null argument");
  int did = 0;
  int rc = gb_pf_xchg_read(pf, nullptr, ess, nullptr, &did);
  if (rc) return rc;
  *did_resample = did;
  if (!did) return GB_OK;
  pf->log_ml_est = pf->h_ctl->log_ml_est;
  pf->wstate = W_ZERO; pf->max_posted = false;
  pf->tiles_valid = false;
  pf->anc_valid = true; pf->parents64_valid = false;
  pf->pending_gather = true;
  pf->ext_used = pf->a_world > 1;
  pf->mh_ok = false;
  return GB_OK;
}

// ---- sharded fast path: slots that stay on this shard keep their ancestors (deferred gather, fused into
//      the next step); rows arriving from other shards are parked in the extension rows of the current
//      trace buffer and the slot's ancestor index points there ----
extern "C" int gb_pf_ext_capacity(const gb_pf_t *pf, int64_t *cap) { REQ(pf && cap, "null argument"); *cap = pf->ext_cap; return GB_OK; }
extern "C" int gb_pf_offspring_local(gb_pf_t *pf, int64_t k_lo, int64_t k_hi) {
  REQ(pf, "null handle");
  REQ(k_lo >= pf->off_lo && k_hi <= pf->off_hi && k_lo <= k_hi, "gb_pf_offspring_local: slots outside the shard's offspring range");
  REQ(k_lo >= pf->pid_offset && k_hi <= pf->pid_offset + pf->n, "gb_pf_offspring_local: slots not owned by this shard");
  return GB_OK;   // the sweep already wrote these ancestors into the shard's `anc` column
}
extern "C" int gb_pf_import_ext(gb_pf_t *pf, int64_t dst_lo, int64_t count, int64_t ext_offset, const void *dev_block) {
  if (pf) pf->mh_ok = false;
  REQ(pf && dev_block, "null argument");
  REQ(dst_lo >= 0 && count >= 0 && dst_lo + count <= pf->n, "gb_pf_import_ext: destination outside the shard");
  if (ext_offset < 0 || ext_offset + count > pf->ext_cap) return fail(GB_ESTATE, "gb_pf_import_ext: extension rows exhausted");
  if (count == 0) return GB_OK;
  if (pf->dtype == GB_F64) {
    auto kern = import_ext_kernel<double>;
    int grid = grid_for(pf, kern, count, kBlock);
    LAUNCH(pf, GB_K_GATHER, kern<<<grid, kBlock, 0, pf->stream>>>((double *)pf->state[pf->cur], pf->ld, pf->S, dst_lo, count, pf->n_pad + ext_offset,
                                                                   dev_block, pf->anc, pf->ext_parents + ext_offset));
  } else {
    auto kern = import_ext_kernel<float>;
    int grid = grid_for(pf, kern, count, kBlock);
    LAUNCH(pf, GB_K_GATHER, kern<<<grid, kBlock, 0, pf->stream>>>((float *)pf->state[pf->cur], pf->ld, pf->S, dst_lo, count, pf->n_pad + ext_offset,
                                                                   dev_block, pf->anc, pf->ext_parents + ext_offset));
  }
  CK(cudaGetLastError());
  pf->ext_used = true;
  return GB_OK;
}
extern "C" int gb_pf_commit_resample_lazy(gb_pf_t *pf, double log_ml_increment) {
  if (pf) pf->mh_ok = false;
  REQ(pf, "null handle");
  pf->log_ml_est += log_ml_increment;       // particle_filter.jl:263
  pf->wstate = W_ZERO;                      // :266 (the array itself is rewritten by the gather / next step)
  pf->tiles_valid = false;
  pf->anc_valid = true;
  pf->parents64_valid = false;
  pf->pending_gather = true;
  int rc = GB_OK;
  if (!pf->lazy) rc = run_gather(pf);
  return rc;
}
// host-side bookkeeping of a resample the CALLER decided (synchronous maybe_resample!) whose offspring were exchanged by
// the device-planned peer stores (gb_pf_async_plan / _sweep / _export + the caller's barrier)
extern "C" int gb_pf_async_commit(gb_pf_t *pf, double log_ml_increment) {
  REQ(pf && pf->a_world > 0, "gb_pf_async_commit: call gb_pf_async_setup first");
  int rc = gb_pf_commit_resample_lazy(pf, log_ml_increment);
  if (pf->a_world > 1) pf->ext_used = true;
  return rc;
}
extern "C" int64_t gb_pf_block_bytes(const gb_pf_t *pf, int64_t count) {
  if (!pf || count < 0) return -1;
  const int64_t raw = count * (8 + (int64_t)pf->S * (int64_t)pf->esz);
  return (raw + 15) / 16 * 16;
}
extern "C" int gb_pf_export_offspring(gb_pf_t *pf, int64_t k_lo, int64_t k_hi, void *dev_block) {
  REQ(pf && dev_block, "null argument");
  REQ(k_lo >= pf->off_lo && k_hi <= pf->off_hi && k_lo <= k_hi, "gb_pf_export_offspring: slots outside the shard's offspring range");
  const long long cnt = k_hi - k_lo;
  if (cnt == 0) return GB_OK;
  // ancestors of owned slots live in the `anc` column, the others in the spill buffer
  const long long own_lo = pf->pid_offset, own_hi = pf->pid_offset + pf->n;
  const bool owned = k_lo >= own_lo && k_hi <= own_hi;
  REQ(owned || k_hi <= own_lo || k_lo >= own_hi, "gb_pf_export_offspring: range straddles the shard boundary");
  const int *anc = owned ? pf->anc + (k_lo - own_lo) : pf->off_anc + (k_lo - pf->off_lo);
  if (pf->dtype == GB_F64) {
    auto kern = export_kernel<double>;
    int grid = grid_for(pf, kern, cnt, kBlock);
    LAUNCH(pf, GB_K_GATHER, kern<<<grid, kBlock, 0, pf->stream>>>((const double *)pf->state[pf->cur], pf->ld, pf->S, anc, cnt, dev_block, pf->pid_offset));
  } else {
    auto kern = export_kernel<float>;
    int grid = grid_for(pf, kern, cnt, kBlock);
    LAUNCH(pf, GB_K_GATHER, kern<<<grid, kBlock, 0, pf->stream>>>((const float *)pf->state[pf->cur], pf->ld, pf->S, anc, cnt, dev_block, pf->pid_offset));
  }
  CK(cudaGetLastError());
  return GB_OK;
}
extern "C" int gb_pf_import_rows(gb_pf_t *pf, int64_t dst_lo, int64_t count, const void *dev_block) {
  if (pf) pf->mh_ok = false;
  REQ(pf && dev_block, "null argument");
  REQ(dst_lo >= 0 && count >= 0 && dst_lo + count <= pf->n, "gb_pf_import_rows: destination outside the shard");
  if (count == 0) return GB_OK;
  if (!pf->parents64) CK(gb_dev_alloc((void **)&pf->parents64, (size_t)pf->ld * 8));
  if (pf->dtype == GB_F64) {
    auto kern = import_kernel<double>;
    int grid = grid_for(pf, kern, count, kBlock);
    LAUNCH(pf, GB_K_GATHER, kern<<<grid, kBlock, 0, pf->stream>>>((double *)pf->state[pf->cur ^ 1], pf->ld, pf->S, dst_lo, count, dev_block, pf->parents64));
  } else {
    auto kern = import_kernel<float>;
    int grid = grid_for(pf, kern, count, kBlock);
    LAUNCH(pf, GB_K_GATHER, kern<<<grid, kBlock, 0, pf->stream>>>((float *)pf->state[pf->cur ^ 1], pf->ld, pf->S, dst_lo, count, dev_block, pf->parents64));
  }
  CK(cudaGetLastError());
  return GB_OK;
}
extern "C" int gb_pf_commit_resample(gb_pf_t *pf, double log_ml_increment) {
  if (pf) pf->mh_ok = false;
  REQ(pf, "null handle");
  { const int urc = unshare_incr(pf); if (urc) return urc; }     // (logw is zeroed below)
  pf->cur ^= 1;
  pf->log_ml_est += log_ml_increment;
  pf->wstate = W_ZERO;
  pf->tiles_valid = false;
  pf->anc_valid = false;
  pf->parents64_valid = true;
  pf->pending_gather = false;
  int rc = pf->dtype == GB_F64 ? fill_dev<double>(pf, (double *)pf->logw, pf->ld, 0.0) : fill_dev<float>(pf, (float *)pf->logw, pf->ld, 0.f);
  return rc;
}

// ------------------------------------------------------------------------------------------
// importance sampling
// ------------------------------------------------------------------------------------------
static int blr_generate(gb_pf *pf, const double *obs, int nobs, bool predrawn) {
  const gb_model &m = pf->model;
  std::vector<double> Xt((size_t)kBlrMax * kBlrMax, 0.0), y(kBlrMax, 0.0);   // [pair k][row j][2]
  for (int j = 0; j < m.nobs; j++)
    for (int d = 0; d < m.D; d++) Xt[((size_t)(d / 2) * kBlrMax + j) * 2 + (d & 1)] = m.params[3 + (size_t)j * m.D + d];
  for (int j = 0; j < nobs; j++) y[j] = obs[j];
  if (!pf->d_tab) CK(gb_dev_alloc((void **)&pf->d_tab, Xt.size() * 8));
  CK(cudaMemcpyAsync(pf->d_tab, Xt.data(), Xt.size() * 8, cudaMemcpyHostToDevice, pf->stream));
  CK(cudaMemcpyToSymbolAsync(c_blr_y, y.data(), y.size() * 8, 0, cudaMemcpyHostToDevice, pf->stream));
  CK(cudaStreamSynchronize(pf->stream));
  BlrArgs a;
  memset(&a, 0, sizeof(a));
  a.state_out = pf->state[pf->cur ^ 1];
  a.logw = pf->logw;
  a.normals = pf->d_normals;
  a.Xt = pf->d_tab;
  a.n = pf->n; a.ld = pf->ld; a.pid_offset = pf->pid_offset;
  a.seed = pf->seed; a.D = m.D; a.nobs = m.nobs; a.sigma = m.params[2];
  a.partials = (double *)pf->partials; a.ticket = pf->ticket; a.stats = pf->d_stats;
#define GO(R, PD)                                                                   \
  {                                                                                 \
    auto kern = blr_generate_kernel<R, PD>;                                         \
    int grid = grid_for(pf, kern, pf->n, 128);                                      \
    LAUNCH(pf, GB_K_STEP, kern<<<grid, 128, 0, pf->stream>>>(a));                   \
  }
  if (pf->dtype == GB_F64) { if (predrawn) GO(double, true) else GO(double, false) }
  else { if (predrawn) GO(float, true) else GO(float, false) }
#undef GO
  CK(cudaGetLastError());
  return GB_OK;
}

extern "C" int gb_importance_sampling(const gb_model_t *m, int64_t n, int dtype, uint64_t seed, const double *obs, int nobs,
                                      int pk, const double *pargs, int npargs, const double *normals, int nn,
                                      const double *uniforms, int nu, double *lognormw_out, double *lml_out, gb_pf_t **pf_out) {
  REQ(m && lml_out, "null argument");
  gb_pf_t *pf = nullptr;
  int rc = gb_pf_create(m, n, n, 0, dtype, seed, &pf);
  if (rc) return rc;
  if ((normals && nn > 0) || (uniforms && nu > 0)) rc = gb_pf_set_noise(pf, normals, nn, uniforms, nu);
  if (!rc) rc = gb_pf_generate(pf, obs, nobs, pk, pargs, npargs);    // importance.jl:28-31, :48-54
  if (!rc) rc = ensure_stats(pf);                                    // :32, :55
  if (rc) { pf_release(pf); return rc; }
  *lml_out = pf->h_stats->lse - log((double)n);                      // :33, :56
  if (lognormw_out) {                                                // :34, :57
    double *d_out = nullptr;
    cudaError_t e = gb_dev_alloc((void **)&d_out, (size_t)n * 8);
    if (e != cudaSuccess) { pf_release(pf); return fail(GB_ECUDA, cudaGetErrorString(e)); }
    if (dtype == GB_F64) {
      auto kern = normalize_kernel<double>;
      int grid = grid_for(pf, kern, n, kBlock);
      LAUNCH(pf, GB_K_MISC, kern<<<grid, kBlock, 0, pf->stream>>>((const double *)pf->logw, n, pf->d_stats, d_out));
    } else {
      auto kern = normalize_kernel<float>;
      int grid = grid_for(pf, kern, n, kBlock);
      LAUNCH(pf, GB_K_MISC, kern<<<grid, kBlock, 0, pf->stream>>>((const float *)pf->logw, n, pf->d_stats, d_out));
    }
    e = cudaGetLastError();
    const int crc = e == cudaSuccess ? staged_d2h(pf->stream, (const double *)d_out, lognormw_out, (size_t)n) : GB_OK;
    gb_dev_free(d_out);
    if (e != cudaSuccess) { pf_release(pf); return fail(GB_ECUDA, cudaGetErrorString(e)); }
    if (crc) { pf_release(pf); return crc; }
  }
  if (pf_out) *pf_out = pf; else pf_release(pf);
  return GB_OK;
}

// ------------------------------------------------------------------------------------------
// raw-array resampling / logsumexp (config 5 microbenchmark)
// ------------------------------------------------------------------------------------------
struct RawWs {       // per-thread cached workspace: a gb_pf shell without particle columns
  gb_pf *pf = nullptr;
  long long n = 0;
  int dtype = -1, device = -1;
};
static thread_local RawWs g_raw;

static int raw_workspace(long long n, int dtype, cudaStream_t stream, gb_pf **out) {
  int rc = require_gpu();
  if (rc) return rc;
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  if (g_raw.pf && g_raw.n == n && g_raw.dtype == dtype && g_raw.device == cur_dev) { g_raw.pf->stream = stream; *out = g_raw.pf; return GB_OK; }
  if (g_raw.pf) { pf_release(g_raw.pf); g_raw.pf = nullptr; }
  REQ(n >= 1 && n < ((int64_t)1 << 31), "raw resampling: 1 <= n < 2^31");
  gb_pf *pf = new gb_pf;
  pf->dtype = dtype; pf->esz = dtype == GB_F64 ? 8 : 4;
  pf->n = pf->n_global = n; pf->ld = pf->n_pad = ((n + kTile - 1) / kTile) * kTile; pf->ntiles = pf->ld / kTile;
  pf->stream = stream;
  int dev = 0;
  cudaGetDevice(&dev);
  pf->device = dev;
  cudaDeviceGetAttribute(&pf->num_sms, cudaDevAttrMultiProcessorCount, dev);
  cudaError_t e = gb_dev_alloc((void **)&pf->partials, sizeof(double2) * kMaxPartials);
  if (e == cudaSuccess) e = alloc_sys_workspace(pf);
  if (e == cudaSuccess) e = gb_dev_alloc((void **)&pf->ticket, 64);
  if (e == cudaSuccess) e = cudaMemset(pf->ticket, 0, 64);
  if (e == cudaSuccess) e = gb_dev_alloc((void **)&pf->d_stats, sizeof(WeightStats));
  if (e == cudaSuccess) e = cudaMemset(pf->d_stats, 0, sizeof(WeightStats));     // (xc = nullptr: not a sharded filter)
  if (e == cudaSuccess) e = gb_host_alloc((void **)&pf->h_stats, sizeof(WeightStats));
  if (e == cudaSuccess) e = gb_dev_alloc((void **)&pf->tile_q, (size_t)(pf->ntiles + 1) * 8);
  if (e == cudaSuccess) e = gb_dev_alloc((void **)&pf->d_plan, sizeof(ResamplePlan));
  if (e == cudaSuccess) e = gb_host_alloc((void **)&pf->h_plan, sizeof(ResamplePlan));
  if (e != cudaSuccess) { pf_release(pf); return fail(GB_ECUDA, std::string("raw workspace: ") + cudaGetErrorString(e)); }
  g_raw.pf = pf; g_raw.n = n; g_raw.dtype = dtype; g_raw.device = dev;
  *out = pf;
  return GB_OK;
}

// max + sums + quantised tile totals of a raw device array -> ws->h_stats
static int raw_stats(gb_pf *ws, const void *dev_logw) {
  ws->wstate = W_DIRTY;
  int rc = ensure_max(ws, dev_logw);
  if (rc) return rc;
  return run_wstats(ws, dev_logw, 0.0, true);
}

// NOTE: dev_logw must be padded to a multiple of 2048 elements (the engine's own arrays are); the
// host-buffer entry below takes care of that, device callers allocate round_up(n, 2048) elements.
extern "C" int gb_resample_device(const void *dev_logw, int64_t n, int dtype, int kind, uint32_t u0, const void *dev_u64,
                                  uint64_t seed, uint32_t step, void *dev_anc, void *cuda_stream) {
  REQ(dev_logw && dev_anc, "null argument");
  REQ(dtype == GB_F64 || dtype == GB_F32, "bad dtype");
  gb_pf *ws = nullptr;
  int rc = raw_workspace(n, dtype, (cudaStream_t)cuda_stream, &ws);
  if (rc) return rc;
  ws->ms_e_async = false;
  if (kind == GB_RESAMPLE_MULTINOMIAL_SORTED && !dev_u64 && (rc = msort_prelaunch_spacings(ws, seed, step))) return rc;
  if ((rc = raw_stats(ws, dev_logw))) return rc;
  const double m = ws->h_stats->m;
  if (!(m > -INFINITY)) {
    if (ws->ms_e_async) { cudaStreamSynchronize(ws->ms_aux); ws->ms_e_async = false; }
    return fail(GB_ESTATE, "gb_resample: all log weights are -Inf or NaN");
  }
  return dtype == GB_F64 ? compute_ancestors<double>(ws, dev_logw, m, kind, u0, (const uint64_t *)dev_u64, seed, step, (int *)dev_anc)
                         : compute_ancestors<float>(ws, dev_logw, m, kind, u0, (const uint64_t *)dev_u64, seed, step, (int *)dev_anc);
}

extern "C" int gb_resample(const double *logw, int64_t n, int dtype, int kind, uint32_t u0, const uint64_t *u64s,
                           uint64_t seed, uint32_t step, int64_t *anc_out) {
  REQ(logw && anc_out && n >= 1, "null argument");
  int rc = require_gpu();
  if (rc) return rc;
  const long long ld = ((n + kTile - 1) / kTile) * kTile;
  const size_t esz = dtype == GB_F64 ? 8 : 4;
  void *d_lw = nullptr;
  int *d_anc = nullptr;
  uint64_t *d_u = nullptr;
  std::vector<float> tmpf;
  std::vector<int> anc32(n);
  cudaError_t e = gb_dev_alloc(&d_lw, (size_t)ld * esz);
  if (e == cudaSuccess) e = cudaMemset(d_lw, 0xFF, (size_t)ld * esz);   // padding = NaN, masked by index anyway
  if (e == cudaSuccess) e = gb_dev_alloc((void **)&d_anc, (size_t)ld * 4);
  if (e == cudaSuccess && u64s) {      // (sorted multinomial: n + 1 spacings instead of n draws)
    const size_t nu = (size_t)n + (kind == GB_RESAMPLE_MULTINOMIAL_SORTED ? 1 : 0);
    e = gb_dev_alloc((void **)&d_u, ((size_t)ld + kTile) * 8);
    if (e == cudaSuccess) e = cudaMemcpy(d_u, u64s, nu * 8, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) {
    if (dtype == GB_F64) e = cudaMemcpy(d_lw, logw, (size_t)n * 8, cudaMemcpyHostToDevice);
    else {
      tmpf.resize(n);
      for (long long i = 0; i < n; i++) tmpf[i] = (float)logw[i];
      e = cudaMemcpy(d_lw, tmpf.data(), (size_t)n * 4, cudaMemcpyHostToDevice);
    }
  }
  rc = GB_OK;
  if (e != cudaSuccess) rc = fail(GB_ECUDA, std::string("gb_resample: ") + cudaGetErrorString(e));
  if (!rc) rc = gb_resample_device(d_lw, n, dtype, kind, u0, d_u, seed, step, d_anc, nullptr);
  if (!rc) {
    e = cudaMemcpy(anc32.data(), d_anc, (size_t)n * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(GB_ECUDA, std::string("gb_resample: ") + cudaGetErrorString(e));
  }
  gb_dev_free(d_lw); gb_dev_free(d_anc); gb_dev_free(d_u);
  if (rc) return rc;
  for (long long i = 0; i < n; i++) anc_out[i] = anc32[i];
  return GB_OK;
}

extern "C" int gb_logsumexp(const double *a, int64_t n, int dtype, double *lse_out, double *ess_out) {
  REQ(a && n >= 1, "null argument");
  gb_pf *ws = nullptr;
  int rc = raw_workspace(n, dtype, nullptr, &ws);
  if (rc) return rc;
  void *d = nullptr;
  const size_t esz = dtype == GB_F64 ? 8 : 4;
  // (the statistics pass loads whole 2048-element tiles and masks by index: the array is padded like the engine's own)
  CK(gb_dev_alloc(&d, (size_t)ws->ld * esz));
  cudaError_t e;
  if (dtype == GB_F64) e = cudaMemcpy(d, a, (size_t)n * 8, cudaMemcpyHostToDevice);
  else {
    std::vector<float> t(n);
    for (long long i = 0; i < n; i++) t[i] = (float)a[i];
    e = cudaMemcpy(d, t.data(), (size_t)n * 4, cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) { gb_dev_free(d); return fail(GB_ECUDA, cudaGetErrorString(e)); }
  rc = raw_stats(ws, d);
  gb_dev_free(d);
  if (rc) return rc;
  if (lse_out) *lse_out = ws->h_stats->lse;
  if (ess_out) *ess_out = ws->h_stats->ess;
  return GB_OK;
}

// tail of enumerative_inference (enumerative.jl:39-46): log_weights[i] = log_weight + log_vol; logsumexp; normalise
extern "C" int gb_enumerative_normalize(const double *log_weights, const double *log_vols, int64_t n, int dtype, double *lognormw_out,
                                        double *log_total_out) {
  REQ(log_weights && n >= 1 && log_total_out, "gb_enumerative_normalize: bad argument");
  std::vector<double> lw(log_weights, log_weights + n);
  if (log_vols)
    for (int64_t i = 0; i < n; i++) lw[i] += log_vols[i];                       // :41
  double lse = 0.0;
  int rc = gb_logsumexp(lw.data(), n, dtype, &lse, nullptr);                    // :43 (device reduction)
  if (rc) return rc;
  *log_total_out = lse;
  if (lognormw_out)
    for (int64_t i = 0; i < n; i++) lognormw_out[i] = lw[i] - lse;              // :44
  return GB_OK;
}

// importance_resampling (importance.jl:77-122), streaming: O(chunk) device memory for any number of proposals
extern "C" int gb_importance_resampling(const gb_model_t *m, int64_t n, int dtype, uint64_t seed, const double *obs, int nobs,
                                        int pk, const double *pargs, int npargs, int64_t chunk, double *row_out, double *lml_out) {
  REQ(m && n >= 1 && row_out && lml_out, "gb_importance_resampling: bad argument");
  if (chunk <= 0) chunk = (int64_t)1 << 20;
  chunk = std::min<int64_t>(std::min<int64_t>(chunk, (int64_t)1 << 22), n);
  gb_pf_t *pf = nullptr;
  int rc = gb_pf_create(m, chunk, chunk, 0, dtype, seed, &pf);
  if (rc) return rc;
  const int S = pf->S;
  std::vector<double> row(S), kept(S);
  double total = -INFINITY;                     // running log total weight (:93, :111)
  uint64_t draw = 0;
  for (int64_t off = 0; off < n && !rc; off += chunk, draw++) {
    const int64_t cnt = std::min<int64_t>(chunk, n - off);
    // the chunk's proposals: global sample ids [off, off + cnt) of the same Philox stream a one-shot run would use
    pf->n = cnt; pf->n_global = cnt; pf->pid_offset = off;
    pf->wstate = W_DIRTY; pf->tiles_valid = false; pf->ws_clean = false;
    if ((rc = gb_pf_generate(pf, obs, nobs, pk, pargs, npargs))) break;     // (:88-92, :106-110 for every proposal of the chunk)
    pf->pid_offset = 0;                          // (indices returned below are chunk-local)
    if ((rc = ensure_stats(pf))) break;
    const double lse = pf->h_stats->lse;
    if (!(lse > -INFINITY)) continue;            // a chunk without weight can never supply the kept trace
    int64_t idx = 0;
    if ((rc = gb_pf_sample_unweighted(pf, 1, (draw << 8) | 0xA5u, &idx, row.data()))) break;
    // fold the chunk into the running total: logsumexp(x1, x2) (inference.jl:8-11), keep the chunk's draw with
    // probability exp(lse - new total) -- the reference's per-proposal Bernoulli swap (:95-103) applied to a block
    const double mx = std::max(total, lse);
    const double nt = mx + log(exp(total - mx) + exp(lse - mx));
    const U4 w = philox_block(seed ^ 0xD1B54A32D192ED03ull, draw, 0u, PK_SAMPLE, 0xFFFFFFu);
    if (u52(w.x, w.y) < exp(lse - nt)) kept = row;
    total = nt;
  }
  pf->n = chunk; pf->n_global = chunk; pf->pid_offset = 0;
  pf_release(pf);
  if (rc) return rc;
  if (!(total > -INFINITY)) return fail(GB_ESTATE, "gb_importance_resampling: every proposal has zero weight");
  for (int s = 0; s < S; s++) row_out[s] = kept[s];
  *lml_out = total - log((double)n);            // :104, :121
  return GB_OK;
}

// ------------------------------------------------------------------------------------------
// noise export + host-side integer helpers
// ------------------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(kBlock) noise_kernel(unsigned long long seed, unsigned int step, long long pid0, long long n,
                                                       int ndraw, int uniform, double *__restrict__ out) {
  __shared__ double s_tab[kMathTabDoubles];
  const MathTables T = stage_math_tables(s_tab);
  __syncthreads();
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock)
    for (int d = 0; d < ndraw; d += 2) {
      R a, b;
      if (uniform) philox_uniform_pair(seed, (unsigned long long)(pid0 + i), step, (unsigned)(d / 2), a, b);
      else philox_normal_pair(seed, (unsigned long long)(pid0 + i), step, (unsigned)(d / 2), T, a, b);
      out[(long long)d * n + i] = (double)a;
      if (d + 1 < ndraw) out[(long long)(d + 1) * n + i] = (double)b;
    }
}
static int export_noise(uint64_t seed, uint32_t step, int64_t pid0, int64_t n, int ndraw, int dtype, int uniform, double *out) {
  REQ(out && n >= 1 && ndraw >= 1, "bad argument");
  int rc = require_gpu();
  if (rc) return rc;
  double *d = nullptr;
  CK(gb_dev_alloc((void **)&d, (size_t)n * ndraw * 8));
  int grid = (int)std::min<long long>((n + kBlock - 1) / kBlock, 148 * 8);
  if (dtype == GB_F64) noise_kernel<double><<<grid, kBlock>>>(seed, step, pid0, n, ndraw, uniform, d);
  else noise_kernel<float><<<grid, kBlock>>>(seed, step, pid0, n, ndraw, uniform, d);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(out, d, (size_t)n * ndraw * 8, cudaMemcpyDeviceToHost);
  gb_dev_free(d);
  if (e != cudaSuccess) return fail(GB_ECUDA, std::string("noise export: ") + cudaGetErrorString(e));
  return GB_OK;
}
extern "C" int gb_philox_normals(uint64_t seed, uint32_t step, int64_t pid0, int64_t n, int ndraw, int dtype, double *out) {
  return export_noise(seed, step, pid0, n, ndraw, dtype, 0, out);
}
extern "C" int gb_philox_uniforms(uint64_t seed, uint32_t step, int64_t pid0, int64_t n, int ndraw, int dtype, double *out) {
  return export_noise(seed, step, pid0, n, ndraw, dtype, 1, out);
}
extern "C" uint32_t gb_philox_resample_u0(uint64_t seed, uint32_t step) { return philox_block(seed, 0, step, PK_RESAMPLE, 0).x; }
extern "C" int gb_philox_multinomial_u64(uint64_t seed, uint32_t step, int64_t k0, int64_t n, uint64_t *out) {
  REQ(out, "null argument");
  for (int64_t i = 0; i < n; i++) {   // integer-only: identical on host and device
    U4 w = philox_block(seed, (uint64_t)(k0 + i), step, PK_MULTI, 0);
    out[i] = ((uint64_t)w.x << 32) | w.y;
  }
  return GB_OK;
}
// the integer exponential spacings GB_RESAMPLE_MULTINOMIAL_SORTED derives from Philox (slots k0 .. k0 + n - 1 of resample
// `step`), computed by the device kernel itself: what the tests hand to the oracle
extern "C" int gb_philox_sorted_spacings(uint64_t seed, uint32_t step, int64_t k0, int64_t n, uint64_t *out) {
  REQ(out && n >= 1 && k0 >= 0, "gb_philox_sorted_spacings: bad argument");
  int rc = require_gpu();
  if (rc) return rc;
  const long long nst = (n + kTile - 1) / kTile;
  unsigned long long *d = nullptr;
  CK(gb_dev_alloc((void **)&d, (size_t)nst * kTile * 8));
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)std::min<long long>(nst, (long long)(sms > 0 ? sms : 1) * 8);
  msort_spacing_kernel<true><<<grid, kBlock>>>(d, n - 1, nst, seed, step, (long long)k0, nullptr);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(out, d, (size_t)n * 8, cudaMemcpyDeviceToHost);
  gb_dev_free(d);
  if (e != cudaSuccess) return fail(GB_ECUDA, std::string("gb_philox_sorted_spacings: ") + cudaGetErrorString(e));
  return GB_OK;
}
extern "C" int gb_quant_bits(int64_t n_global) { return quant_bits(n_global); }

// host-side evaluation of the exact-integer helpers the kernels use (so the CPU test-suite can check
// them against Python big integers without a GPU); not part of the product path
extern "C" int64_t gb_selftest_slots_below(uint64_t C, int64_t nslots, uint64_t q_total, uint32_t u0) {
  SweepParams sp;
  sp.q_total = q_total; sp.q_inv = 1.0 / (double)q_total; sp.dscale = (uint64_t)nslots << 32; sp.u0 = u0;
  return slots_below(C, sp);
}
extern "C" int64_t gb_selftest_slots_below_residual(uint64_t chi, uint64_t clo, uint64_t q_total, uint32_t u0) {
  SweepParams sp;
  sp.q_total = q_total; sp.q_inv = 1.0 / (double)q_total; sp.dscale = 0; sp.u0 = u0;
  return slots_below_residual(chi, clo, sp);
}
extern "C" uint64_t gb_selftest_div128(uint64_t ahi, uint64_t alo, uint64_t d) { return div128_64(ahi, alo, d, 1.0 / (double)d); }
extern "C" uint64_t gb_selftest_quantize(double x, int bits) { return quantize_k(x, bits, h_wexp_tab); }
// fn: 0 log(a), 1 sin(2 pi a), 2 cos(2 pi a), 3 atan2(a, b), 4 exp(a), 5 u52_fast(hi = a, lo = b)
extern "C" double gb_selftest_math(int fn, double a, double b) {
  MathTables T = host_math_tables();
  double s, c;
  switch (fn) {
    case 0: return gb_log(a, T);
    case 1: gb_sincos2pi(a, s, c); return s;
    case 2: gb_sincos2pi(a, s, c); return c;
    case 3: return gb_atan2(a, b, T);
    case 4: return gb_exp(a);
    case 5: return u52_fast((uint32_t)a, (uint32_t)b);
  }
  return NAN;
}
extern "C" void gb_selftest_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  U4 r = philox4x32_10(U4{ctr[0], ctr[1], ctr[2], ctr[3]}, key[0], key[1]);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// ------------------------------------------------------------------------------------------
// batched distribution kernels
// ------------------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(kBlock) dist_logpdf_kernel(int dist, long long n, const double *__restrict__ x, const double *__restrict__ a,
                                                             int as, const double *__restrict__ b, int bs, int len, double logdet,
                                                             double *__restrict__ out) {
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
    R r;
    switch (dist) {
      case GB_DIST_NORMAL: r = logpdf_normal<R>((R)x[i], (R)a[i * as], (R)b[i * bs]); break;
      case GB_DIST_GAMMA: r = logpdf_gamma<R>((R)x[i], (R)a[i * as], (R)b[i * bs]); break;
      case GB_DIST_BERNOULLI: r = logpdf_bernoulli<R>(x[i] != 0.0, (R)a[i * as]); break;
      case GB_DIST_CATEGORICAL: r = logpdf_categorical<R>((long long)x[i], a, len); break;
      case GB_DIST_BROADCASTED_NORMAL: r = logpdf_normal<R>((R)x[i], (R)a[i * as], (R)b[i * bs]); break;   // the host sums (normal.jl:66-67)
      case GB_DIST_DIRICHLET: r = (R)logpdf_dirichlet(x + i * len, a, len); break;
      case GB_DIST_PIECEWISE_UNIFORM: r = (R)logpdf_piecewise_uniform(x[i], a, b, len); break;
      case GB_DIST_BETA_UNIFORM: r = (R)logpdf_beta_uniform(x[i], a[i * as], b[2 * i * bs], b[2 * i * bs + 1]); break;
      case GB_DIST_UNIFORM: case GB_DIST_EXPONENTIAL: case GB_DIST_LAPLACE: case GB_DIST_CAUCHY: case GB_DIST_INV_GAMMA:
      case GB_DIST_BETA: case GB_DIST_POISSON: case GB_DIST_GEOMETRIC: case GB_DIST_BINOM: case GB_DIST_NEG_BINOM:
      case GB_DIST_UNIFORM_DISCRETE:
        r = (R)logpdf_scalar(dist, x[i], a[i * as], b ? b[i * bs] : 0.0);
        break;
      default: {
        R xv[kBlrMax];
        for (int k = 0; k < len; k++) xv[k] = (R)x[i * len + k];
        r = logpdf_mvnormal_chol<R, kBlrMax>(xv, a, b, logdet, len);
      }
    }
    out[i] = (double)r;
  }
}
template <typename R>
__global__ void __launch_bounds__(kBlock) dist_random_kernel(int dist, long long n, const double *__restrict__ a, int as,
                                                             const double *__restrict__ b, int bs, int len, unsigned long long seed,
                                                             unsigned int step, double *__restrict__ out) {
  __shared__ double s_tab[kMathTabDoubles];
  const MathTables T = stage_math_tables(s_tab);
  __syncthreads();
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
    const unsigned long long pid = (unsigned long long)i;
    switch (dist) {
      case GB_DIST_NORMAL: {
        R e0, e1;
        philox_normal_pair(seed, pid, step, 0u, T, e0, e1);
        out[i] = (double)random_normal<R>((R)a[i * as], (R)b[i * bs], e0);
        break;
      }
      case GB_DIST_GAMMA: out[i] = (double)random_gamma_philox<R>((R)a[i * as], (R)b[i * bs], seed, pid, step); break;
      case GB_DIST_BERNOULLI: {
        R u0, u1;
        philox_uniform_pair(seed, pid, step, 0u, u0, u1);
        out[i] = random_bernoulli<R>((R)a[i * as], u0) ? 1.0 : 0.0;
        break;
      }
      case GB_DIST_CATEGORICAL: {
        R u0, u1;
        philox_uniform_pair(seed, pid, step, 0u, u0, u1);
        out[i] = (double)random_categorical<R>(a, len, u0);
        break;
      }
      case GB_DIST_UNIFORM: case GB_DIST_EXPONENTIAL: case GB_DIST_LAPLACE: case GB_DIST_CAUCHY: case GB_DIST_GEOMETRIC:
      case GB_DIST_UNIFORM_DISCRETE: {   // inverse-CDF samplers on one 52-bit uniform in (0,1)
        double u0, u1;
        philox_uniform_pair(seed, pid, step, 0u, u0, u1);
        const double v = random_invcdf(dist, a[i * as], b ? b[i * bs] : 0.0, u0);
        out[i] = v;
        break;
      }
      case GB_DIST_BROADCASTED_NORMAL: {      // normal.jl:123-126: mu .+ std .* randn(shape)
        R e0, e1;
        philox_normal_pair(seed, pid, step, 0u, T, e0, e1);
        out[i] = (double)random_normal<R>((R)a[i * as], (R)b[i * bs], e0);
        break;
      }
      case GB_DIST_POISSON: case GB_DIST_BINOM: case GB_DIST_NEG_BINOM: {   // exact inversion on one 52-bit uniform
        double u0, u1;
        philox_uniform_pair(seed, pid, step, 0u, u0, u1);
        out[i] = dist == GB_DIST_POISSON ? random_poisson(a[i * as], u0)
                 : dist == GB_DIST_BINOM ? random_binom(a[i * as], b[i * bs], u0) : random_neg_binom(a[i * as], b[i * bs], u0);
        break;
      }
      case GB_DIST_DIRICHLET: {               // dirichlet.jl:33-35: independent gamma(alpha_j, 1) draws, normalised
        double tot = 0.0;
        for (int j = 0; j < len; j++) {
          const double gj = random_gamma_philox<double>(a[j], 1.0, seed ^ (0x9E3779B97F4A7C15ull * (unsigned long long)(j + 1)), pid, step);
          out[i * len + j] = gj;
          tot += gj;
        }
        for (int j = 0; j < len; j++) out[i * len + j] /= tot;
        break;
      }
      case GB_DIST_PIECEWISE_UNIFORM: {       // piecewise_uniform.jl:48-52: bin = categorical(probs); uniform on the bin
        double u0, u1;
        philox_uniform_pair(seed, pid, step, 0u, u0, u1);
        const long long bin = random_categorical<double>(b, len, u0) - 1;
        out[i] = u1 * (a[bin + 1] - a[bin]) + a[bin];      // uniform_continuous.jl:21-23
        break;
      }
      case GB_DIST_BETA_UNIFORM: {            // beta_uniform.jl:36-42
        double u0, u1;
        philox_uniform_pair(seed, pid, step, 0u, u0, u1);
        if (u0 < a[i * as]) {
          const double g1 = random_gamma_philox<double>(b[2 * i * bs], 1.0, seed, pid, step);
          const double g2 = random_gamma_philox<double>(b[2 * i * bs + 1], 1.0, seed ^ 0x9E3779B97F4A7C15ull, pid, step);
          out[i] = g1 / (g1 + g2);
        } else {
          out[i] = u1;
        }
        break;
      }
      case GB_DIST_INV_GAMMA: out[i] = b[i * bs] / random_gamma_philox<double>(a[i * as], 1.0, seed, pid, step); break;
      case GB_DIST_BETA: {
        const double g1 = random_gamma_philox<double>(a[i * as], 1.0, seed, pid, step);
        const double g2 = random_gamma_philox<double>(b[i * bs], 1.0, seed ^ 0x9E3779B97F4A7C15ull, pid, step);
        out[i] = g1 / (g1 + g2);
        break;
      }
      default: {   // mvnormal.jl:30-33: mu + L * randn(k)
        R eps[kBlrMax];
        for (int k = 0; k < len; k += 2) {
          R e0, e1;
          philox_normal_pair(seed, pid, step, (unsigned)(k / 2), T, e0, e1);
          eps[k] = e0;
          if (k + 1 < len) eps[k + 1] = e1;
        }
        for (int r = 0; r < len; r++) {
          R s = (R)a[r];
          for (int c = 0; c <= r; c++) s += (R)b[r * len + c] * eps[c];
          out[i * len + r] = (double)s;
        }
      }
    }
  }
}

// lower Cholesky factor (row-major) of a symmetric positive definite matrix; returns logdet
static bool host_cholesky(const double *cov, int k, std::vector<double> &L, double *logdet) {
  L.assign((size_t)k * k, 0.0);
  double ld = 0.0;
  for (int j = 0; j < k; j++) {
    double d = cov[j * k + j];
    for (int p = 0; p < j; p++) d -= L[j * k + p] * L[j * k + p];
    if (!(d > 0.0)) return false;
    d = sqrt(d);
    L[j * k + j] = d;
    ld += 2.0 * log(d);
    for (int i = j + 1; i < k; i++) {
      double s = cov[i * k + j];
      for (int p = 0; p < j; p++) s -= L[i * k + p] * L[j * k + p];
      L[i * k + j] = s / d;
    }
  }
  *logdet = ld;
  return true;
}

static int dist_common(int dist, int dtype, int64_t n, const double *x, const double *a, int as, const double *b, int bs, int len,
                       bool random, uint64_t seed, uint32_t step, double *out) {
  REQ(out && n >= 1, "bad argument");
  REQ(dist >= GB_DIST_NORMAL && dist <= GB_DIST_BETA_UNIFORM, "unknown distribution");
  int rc = require_gpu();
  if (rc) return rc;
  const bool vec = dist == GB_DIST_MVNORMAL;
  const bool dirichlet = dist == GB_DIST_DIRICHLET, pwu = dist == GB_DIST_PIECEWISE_UNIFORM, bu = dist == GB_DIST_BETA_UNIFORM;
  if (dist == GB_DIST_CATEGORICAL || vec || dirichlet || pwu) REQ(len >= 1 && (!vec || len <= kBlrMax), "len out of range (mvnormal: k <= 64)");
  if (pwu) REQ(b, "piecewise_uniform needs bounds (a, len + 1 values) and probs (b, len values)");
  if (bu) REQ(b, "beta_uniform needs b = (alpha, beta)");
  const bool rowwise = vec || dirichlet;                       // x / the draws are [n][len]
  size_t nx = (size_t)n * (rowwise ? len : 1);
  size_t na = (dist == GB_DIST_CATEGORICAL || rowwise) ? (size_t)len : (pwu ? (size_t)len + 1 : (as ? (size_t)n : 1));
  size_t nb = vec ? (size_t)len * len : (pwu ? (size_t)len : (b ? (bs ? (size_t)n : 1) * (bu ? 2 : 1) : 0));
  std::vector<double> L;
  double logdet = 0.0;
  const double *bsrc = b;
  if (vec) {
    REQ(b, "mvnormal needs a covariance");
    if (!host_cholesky(b, len, L, &logdet)) return fail(GB_EINVAL, "mvnormal: covariance is not positive definite");
    bsrc = L.data();
  }
  if (pwu)
    for (int j = 0; j < len; j++) REQ(a[j] < a[j + 1], "piecewise_uniform: bounds must be strictly increasing");
  double *dx = nullptr, *da = nullptr, *db = nullptr, *dout = nullptr;
  cudaError_t e = cudaSuccess;
  size_t nout = random ? nx : (size_t)n;
  if (!random) { e = gb_dev_alloc((void **)&dx, nx * 8); if (e == cudaSuccess) e = cudaMemcpy(dx, x, nx * 8, cudaMemcpyHostToDevice); }
  if (e == cudaSuccess) e = gb_dev_alloc((void **)&da, na * 8);
  if (e == cudaSuccess) e = cudaMemcpy(da, a, na * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && nb) { e = gb_dev_alloc((void **)&db, nb * 8); if (e == cudaSuccess) e = cudaMemcpy(db, bsrc, nb * 8, cudaMemcpyHostToDevice); }
  if (e == cudaSuccess) e = gb_dev_alloc((void **)&dout, nout * 8);
  if (e == cudaSuccess) {
    int grid = (int)std::min<long long>((n + kBlock - 1) / kBlock, 148 * 8);
    const bool shared_a = dist == GB_DIST_CATEGORICAL || rowwise || pwu;
    const int as_ = shared_a ? 0 : (as ? 1 : 0), bs_ = (vec || pwu) ? 0 : (bs ? 1 : 0);
    const double *dbb = db ? db : da;   // bernoulli / categorical have no second parameter
    if (random) {
      if (dtype == GB_F64) dist_random_kernel<double><<<grid, kBlock>>>(dist, n, da, as_, dbb, bs_, len, seed, step, dout);
      else dist_random_kernel<float><<<grid, kBlock>>>(dist, n, da, as_, dbb, bs_, len, seed, step, dout);
    } else {
      if (dtype == GB_F64) dist_logpdf_kernel<double><<<grid, kBlock>>>(dist, n, dx, da, as_, dbb, bs_, len, logdet, dout);
      else dist_logpdf_kernel<float><<<grid, kBlock>>>(dist, n, dx, da, as_, dbb, bs_, len, logdet, dout);
    }
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) {
    if (!random && dist == GB_DIST_BROADCASTED_NORMAL) {       // normal.jl:66-67: ONE number, the sum over the broadcast shape
      std::vector<double> tmp(nout);
      e = cudaMemcpy(tmp.data(), dout, nout * 8, cudaMemcpyDeviceToHost);
      double acc = 0.0;
      for (size_t i = 0; i < nout; i++) acc += tmp[i];
      out[0] = acc;
    } else {
      e = cudaMemcpy(out, dout, nout * 8, cudaMemcpyDeviceToHost);
    }
  }
  gb_dev_free(dx); gb_dev_free(da); gb_dev_free(db); gb_dev_free(dout);
  if (e != cudaSuccess) return fail(GB_ECUDA, std::string("gb_dist: ") + cudaGetErrorString(e));
  return GB_OK;
}
extern "C" int gb_dist_logpdf(int dist, int dtype, int64_t n, const double *x, const double *a, int as, const double *b, int bs,
                              int len, double *out) {
  REQ(x && a, "null argument");
  return dist_common(dist, dtype, n, x, a, as, b, bs, len, false, 0, 0, out);
}
extern "C" int gb_dist_random(int dist, int dtype, int64_t n, const double *a, int as, const double *b, int bs, int len,
                              uint64_t seed, uint32_t step, double *out) {
  REQ(a, "null argument");
  return dist_common(dist, dtype, n, nullptr, a, as, b, bs, len, true, seed, step, out);
}

// out[i] -= shift over a large host array (importance.jl:34,57 on the gathered log weights)
static void host_shift(double *a, size_t n, double shift) {
  const unsigned hw = std::thread::hardware_concurrency();
  const int T = (n < (1u << 18)) ? 1 : (int)std::max(1u, std::min(8u, hw ? hw : 1u));
  auto work = [=](int t) { for (size_t i = n * t / T, hi = n * (t + 1) / T; i < hi; i++) a[i] -= shift; };
  std::vector<std::thread> th;
  for (int t = 1; t < T; t++) th.emplace_back(work, t);
  work(0);
  for (auto &x : th) x.join();
}

#include "spf.inc"
