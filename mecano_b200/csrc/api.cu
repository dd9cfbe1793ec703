// api.cu -- the C ABI declared in include/mecano_b200.h.  Host-side plumbing only: argument checks,
// handle lifetime, kernel planning, and the chunked host<->device pipeline behind the *_host entry
// points.  There is deliberately no CPU compute path in this library: every compute entry point ends
// in a kernel launch or returns an error.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/mecano_b200.h"
#include "flatten.h"
#include "jit.h"
#include "kernels.h"
#include "specialize.h"

#define MB_STAGGER_NS_DEFAULT 0
#define MB_ABA_DISCARD_DEFAULT 1 // profiles/r04b_aba_records.md: 1.713 -> 1.696 ms, DRAM writes 2.46 -> 1.82 GB per 2^20 H37 states
#define MB_MAX_SLOTS 4
#define MB_COUNTERS 256

struct mecano_b200_handle
{
   int device = 0;
   mb::FlatTree tree;
   double *d_consts = nullptr;
   uint16_t *d_zero = nullptr; // CRBA: structurally zero entries
   MbProgram *d_prog = nullptr; // [3] device copies of the traversal programs (warp-per-state kernels)
   std::vector<std::pair<int, int>> nonzero_runs; // CRBA: (first entry, count) runs of mass-matrix entries that are not structurally zero
   unsigned *d_counters = nullptr; // work counters of the persistent launches (kernels.cu: launch_thread_kernel): a ring of MB_COUNTERS,
   unsigned counter_seq = 0;       // a fresh one per launch, so that launches in flight on different streams never share one
   double *d_zero_row = nullptr; // one row of zeros: stands in for qd / qdd when RNEA ignores velocities / accelerations
   size_t zero_row_doubles = 0;
   double *d_scratch = nullptr; // joint efforts nobody asked for (centroidal convective term = one RNEA launch)
   size_t scratch_doubles = 0;
   double *d_ws[1 + MB_MAX_SLOTS] = {}; // ABA pass-two records: [0] device entry points, [1 + s] host-pipeline slot s
   size_t ws_doubles[1 + MB_MAX_SLOTS] = {};
   mb::SpecKernel spec[MB_NUM_ALGOS];                        // tree-specialised kernels (mecano_b200_specialize), per algorithm
   double gravity[3] = {0.0, 0.0, 0.0}; // Mecano calculators start with zero gravity until setGravitationalAcceleration
   mb::LaunchPlan plan[MB_NUM_ALGOS];
   bool fp32 = false; // mecano_b200_set_precision
   int grid_limit[MB_NUM_ALGOS] = {0, 0, 0, 0}; // mecano_b200_set_grid_limit: cap on the persistent grids (0 = whole device)
   int variant = MECANO_B200_VARIANT_AUTO;
   int max_children = 1, max_ndof = 1, sm_count = 148;
   int n_accel_source = 0;           // joints in ACCELERATION_SOURCE mode (mecano_b200_set_joint_source_modes)
   std::vector<std::pair<int, int>> effort_dof_runs; // (first DoF row, count) runs of DoF rows whose joints are EFFORT_SOURCE
   bool warp_ok = false;             // the body-parallel variant serves this tree (one-DoF / SixDoF joints; <= 32 bodies: a warp per state, else a team of warps)
   bool has_3dof = false;            // the tree has spherical / planar joints (generic thread-per-state kernels only)
   int64_t warp_below[MB_NUM_ALGOS] = {0, 0, 0, 0}; // AUTO: batches smaller than this run warp-per-state
   std::string error;
   // host pipeline (lazy)
   cudaStream_t streams[MB_MAX_SLOTS] = {};
   cudaEvent_t done[MB_MAX_SLOTS] = {};
   double *stage[MB_MAX_SLOTS] = {};
   size_t stage_doubles = 0;
   int n_slots = 3; // slots (stream + staging buffer) of the host pipeline (profiles/r04b_host_pipe.jsonl: 3 x 128 MB is 15 % faster than 2 x 64 MB)
   std::mutex mu;
};

namespace
{
std::string g_create_error;
std::mutex g_create_mu;

int fail(mecano_b200_handle *h, int code, const std::string &msg)
{
   if (h)
      h->error = msg;
   else
   {
      std::lock_guard<std::mutex> lk(g_create_mu);
      g_create_error = msg;
   }
   return code;
}

int cuda_fail(mecano_b200_handle *h, cudaError_t e, const char *what)
{
   return fail(h, (int)e > 0 ? (int)e : 1, std::string(what) + ": " + cudaGetErrorString(e));
}

#define MB_CUDA(h, call)                                  \
   do                                                     \
   {                                                      \
      cudaError_t e_ = (call);                            \
      if (e_ != cudaSuccess) return cuda_fail(h, e_, #call); \
   } while (0)

// Makes the handle's device current for the duration of an entry point and restores the caller's afterwards (a binding that
// keeps its own notion of the current device -- torch, a JVM thread pool -- must not find it changed behind its back).
struct DeviceGuard
{
   int prev = -1;
   cudaError_t err = cudaSuccess;
   explicit DeviceGuard(int device)
   {
      if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
      if (prev != device) err = cudaSetDevice(device);
      else prev = -1;
   }
   ~DeviceGuard()
   {
      if (prev >= 0) cudaSetDevice(prev);
   }
};
#define MB_ON_DEVICE(h)                                                            \
   DeviceGuard device_guard_((h)->device);                                         \
   if (device_guard_.err != cudaSuccess) return cuda_fail(h, device_guard_.err, "cudaSetDevice")

int check_batch(mecano_b200_handle *h, int64_t n, int64_t ld)
{
   if (!h) return MECANO_B200_ERR_INVALID_ARGUMENT;
   if (n < 0) return fail(h, MECANO_B200_ERR_SHAPE, "n_states must be >= 0");
   if (ld < n) return fail(h, MECANO_B200_ERR_SHAPE, "ld must be >= n_states");
   return MECANO_B200_OK;
}

// optional buffers of one launch; any of them routes the call to the generic thread-per-state kernel
struct RunOpts
{
   int ws_slot = 0;                                      // ABA workspace: 0 device entry points, 1 / 2 the host-pipeline slots
   double *body_acc = nullptr, *joint_wrench = nullptr; // RNEA by-products
   const double *x2 = nullptr;                           // ABA: accelerations of the ACCELERATION_SOURCE joints
   double *cmm = nullptr, *com = nullptr;                // CRBA: centroidal momentum matrix (root frame), (mass * CoM, mass)
   double *root_wrench = nullptr;                        // RNEA: wrench at the root (root frame)
   double *cor = nullptr;                                // MB_CORIOLIS: the Coriolis matrix
   bool zero_gravity = false;
   int64_t variant_n = 0; // batch size VARIANT_AUTO decides on (0: n of this launch); the host pipelines pass the size of the whole call so
                          // that every chunk and every device slice of one call runs the same kernel (bit-identical results)
};

int run(mecano_b200_handle *h, int algo, int64_t n, int64_t ld, const double *q, const double *qd, const double *x, const double *fext,
        double *out, uint32_t flags, cudaStream_t stream, const RunOpts &opt = RunOpts())
{
   const int ws_slot = opt.ws_slot;
   double *const body_acc = opt.body_acc, *const joint_wrench = opt.joint_wrench;
   const double *x2 = opt.x2;
   if (n > (int64_t)1 << 28 || ld > (int64_t)1 << 28)
      return fail(h, MECANO_B200_ERR_TOO_LARGE, "more than 2^28 states (or ld > 2^28) in one call: split the batch");
   if (algo == MB_ABA && h->n_accel_source > 0 && !x2)
      return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "joints in ACCELERATION_SOURCE mode need their accelerations: call mecano_b200_aba_sources");
   if (h->n_accel_source == 0)
      x2 = nullptr; // nothing reads it: the plain kernels serve the call
   const bool packed = algo == MB_CRBA && (flags & MECANO_B200_CRBA_PACKED);
   if (packed && ((flags & (MECANO_B200_CRBA_STATE_MAJOR | MECANO_B200_CRBA_ZEROS_PRESENT)) || opt.cmm || opt.com))
      return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "the packed mass-matrix layout does not combine with STATE_MAJOR / ZEROS_PRESENT or the centroidal by-products");
   if (h->fp32)
   {
      // the optional fp32 variant: plain calls on the thread-per-state kernels only, never a silent fp64 substitute
      if (algo > MB_CRBA || packed || fext || opt.body_acc || opt.joint_wrench || x2 || opt.cmm || opt.com || opt.root_wrench || flags != 0)
         return fail(h, MECANO_B200_ERR_UNSUPPORTED_TOPOLOGY, "the fp32 variant covers plain RNEA / ABA / CRBA calls only (no external wrenches, flags, by-products)");
      if (!h->plan[algo].fp32_ok || h->has_3dof)
         return fail(h, MECANO_B200_ERR_UNSUPPORTED_TOPOLOGY, "the fp32 variant is compiled for the launch configuration of humanoid-sized trees of one-DoF and SixDoF joints only");
   }
   // thread- or warp-per-state: explicit choice, else by batch size (a warp per state fills the machine from a few hundred
   // states on; a thread per state needs tens of thousands but then has 10-30x the throughput)
   // calls with by-product buffers (mecano_b200_rnea_full) always run the generic thread-per-state kernel
   // (so does the packed mass-matrix layout: its row numbering follows the thread-per-state traversal)
   const bool byprod = body_acc || joint_wrench || x2 || opt.cmm || opt.com || opt.root_wrench || algo == MB_CORIOLIS || h->fp32 || packed;
   const int64_t vn = opt.variant_n > 0 ? opt.variant_n : n;
   const bool use_warp = !byprod && (h->variant == MECANO_B200_VARIANT_WARP || (h->variant == MECANO_B200_VARIANT_AUTO && h->warp_ok && vn < h->warp_below[algo]));
   if (use_warp)
   {
      if (!h->warp_ok)
         return fail(h, MECANO_B200_ERR_UNSUPPORTED_TOPOLOGY, "the warp-per-state variant handles trees of one-DoF and SixDoF joints only");
      mb::KernelArgs wa;
      wa.q = q; wa.qd = qd; wa.x = x; wa.fext = fext; wa.out = out;
      wa.body_acc = wa.joint_wrench = nullptr;
      wa.x2 = nullptr;
      wa.cmm = wa.com = wa.root_wrench = wa.cor = nullptr;
      wa.work_counter = nullptr;
      wa.fp32 = 0;
      wa.consts = h->d_consts;
      wa.ws = nullptr;
      wa.ws_ld = 0;
      wa.zero_entries = h->d_zero;
      wa.n_zero = (algo == MB_CRBA && (flags & MECANO_B200_CRBA_ZEROS_PRESENT)) ? 0 : (int32_t)h->tree.zero_entries.size();
      wa.n = n; wa.ld = ld;
      wa.ld_qd = wa.ld_x = ld;
      wa.grav[0] = h->gravity[0]; wa.grav[1] = h->gravity[1]; wa.grav[2] = h->gravity[2];
      wa.flags = flags;
      wa.nv = h->tree.nv;
      // one lane per body: a warp per state up to 32 bodies, a team of two to four warps beyond (team_kernels.cu)
      if (h->tree.prog[algo].nb <= 32)
         MB_CUDA(h, mb::launch_warp_kernel(algo, h->d_prog + algo, wa, h->max_children, h->max_ndof, h->sm_count, stream));
      else
         MB_CUDA(h, mb::launch_team_kernel(algo, h->d_prog + algo, h->tree.prog[algo].nb, wa, h->max_children, h->max_ndof, h->sm_count, stream));
      return MECANO_B200_OK;
   }
   // the tree-specialised kernel covers the common call (no external wrenches, default flags / layout); everything else
   // runs the generic kernels
   const mb::SpecKernel &sk = h->spec[algo];
   const bool use_spec = sk.ready() && !fext && !byprod && flags == 0;
   if (algo == MB_ABA)
   {
      const size_t need = use_spec ? (size_t)sk.grid * sk.opt.block * (size_t)std::max(h->tree.prog[MB_ABA].rec_doubles, 1) : h->plan[MB_ABA].ws_doubles;
      if (h->ws_doubles[ws_slot] < need)
      {
         if (h->d_ws[ws_slot])
         {
            MB_CUDA(h, cudaStreamSynchronize(stream));
            MB_CUDA(h, cudaFree(h->d_ws[ws_slot]));
            h->d_ws[ws_slot] = nullptr;
            h->ws_doubles[ws_slot] = 0;
         }
         MB_CUDA(h, cudaMalloc(&h->d_ws[ws_slot], need * sizeof(double)));
         h->ws_doubles[ws_slot] = need;
      }
   }
   mb::KernelArgs a;
   a.q = q; a.qd = qd; a.x = x; a.fext = fext; a.out = out;
   a.body_acc = body_acc; a.joint_wrench = joint_wrench;
   a.x2 = x2;
   a.cmm = opt.cmm; a.com = opt.com; a.root_wrench = opt.root_wrench;
   a.cor = opt.cor;
   a.fp32 = h->fp32 ? 1 : 0;
   a.consts = h->d_consts;
   a.ws = h->d_ws[ws_slot];
   a.ws_ld = 0;
   a.zero_entries = h->d_zero;
   a.n_zero = (algo == MB_CRBA && (flags & (MECANO_B200_CRBA_ZEROS_PRESENT | MECANO_B200_CRBA_PACKED))) ? 0 : (int32_t)h->tree.zero_entries.size();
   a.n = n; a.ld = ld;
   a.ld_qd = a.ld_x = ld;
   if (algo == MB_RNEA && (flags & (MECANO_B200_RNEA_NO_CORIOLIS | MECANO_B200_RNEA_NO_ACCELERATIONS)))
   {
      // setConsiderCoriolisAndCentrifugalForces(false) / setConsiderJointAccelerations(false) (InverseDynamicsCalculator.java:
      // 291-306) = the same recursion with zero joint velocities / accelerations: one row of zeros with stride 0 replaces the
      // input, so the kernels carry no flag tests
      if (h->zero_row_doubles < (size_t)n)
      {
         if (h->d_zero_row)
         {
            MB_CUDA(h, cudaDeviceSynchronize());
            MB_CUDA(h, cudaFree(h->d_zero_row));
            h->d_zero_row = nullptr;
            h->zero_row_doubles = 0;
         }
         MB_CUDA(h, cudaMalloc(&h->d_zero_row, (size_t)n * sizeof(double)));
         MB_CUDA(h, cudaMemset(h->d_zero_row, 0, (size_t)n * sizeof(double)));
         h->zero_row_doubles = (size_t)n;
      }
      if (flags & MECANO_B200_RNEA_NO_CORIOLIS) { a.qd = h->d_zero_row; a.ld_qd = 0; }
      if (flags & MECANO_B200_RNEA_NO_ACCELERATIONS) { a.x = h->d_zero_row; a.ld_x = 0; }
   }
   a.grav[0] = h->gravity[0]; a.grav[1] = h->gravity[1]; a.grav[2] = h->gravity[2];
   if (opt.zero_gravity)
      a.grav[0] = a.grav[1] = a.grav[2] = 0.0;
   a.flags = flags;
   if (algo == MB_ABA)
   {
      // pass three discards each record from L2 once read (gpu_ctx.cuh: rec_discard); MECANO_B200_ABA_DISCARD=0 / 1 overrides
      static const int discard = [] { const char *e = getenv("MECANO_B200_ABA_DISCARD"); return e ? atoi(e) : MB_ABA_DISCARD_DEFAULT; }();
      if (discard) a.flags |= MB_KFLAG_ABA_DISCARD;
   }
   a.nv = h->tree.nv;
   a.work_counter = (h->d_counters && !use_spec) ? h->d_counters + 2 * (h->counter_seq++ % MB_COUNTERS) : nullptr;
   {
      // MECANO_B200_STAGGER_NS overrides the start offset between the warps of a scheduler (gpu_ctx.cuh: thread_block_run)
      static const int stagger = [] { const char *e = getenv("MECANO_B200_STAGGER_NS"); return e ? atoi(e) : MB_STAGGER_NS_DEFAULT; }();
      a.stagger_ns = stagger;
   }
   if (use_spec)
   {
      const long long ntiles = (n + sk.opt.block - 1) / sk.opt.block;
      // ABA: persistent grid (workspace column per resident thread); RNEA / CRBA: one block per tile
      const unsigned grid = (algo == MB_ABA || sk.opt.tm > 0) ? (unsigned)std::min<long long>(ntiles, sk.grid) : (unsigned)ntiles;
      a.ws_ld = (long long)sk.grid * sk.opt.block;
      MB_CUDA(h, mb::spec_launch(sk, a, grid, stream));
      return MECANO_B200_OK;
   }
   if (h->grid_limit[algo] > 0 && h->grid_limit[algo] < h->plan[algo].grid)
   {
      mb::LaunchPlan p = h->plan[algo];
      p.grid = h->grid_limit[algo]; // the workspace keeps one column per thread of the full grid: enough for any smaller one
      MB_CUDA(h, mb::launch_thread_kernel(algo, h->tree.prog[algo], a, p, stream));
      return MECANO_B200_OK;
   }
   MB_CUDA(h, mb::launch_thread_kernel(algo, h->tree.prog[algo], a, h->plan[algo], stream));
   return MECANO_B200_OK;
}

// ForwardDynamicsCalculator.compute(tau, qdd_in) with per-joint source modes (ForwardDynamicsCalculator.java:508-520): passes one to
// three in the ABA kernel; pass four (:1315-1363, the efforts of the ACCELERATION_SOURCE joints from the joint wrenches) is the upward
// sweep of inverse dynamics on the accelerations just computed, i.e. one RNEA launch, after which the rows of the EFFORT_SOURCE
// joints are restored to the caller's input as getJointTauMatrix() returns them (:566-590).
int run_aba_sources(mecano_b200_handle *h, int64_t n, int64_t ld, const double *q, const double *qd, const double *tau, const double *qdd_in,
                    const double *fext, double *qdd, double *tau_out, cudaStream_t stream, int ws_slot)
{
   if (h->n_accel_source > 0 && !qdd_in)
      return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "NULL buffer (qdd_in) with joints in ACCELERATION_SOURCE mode");
   RunOpts opt;
   opt.ws_slot = ws_slot;
   opt.x2 = qdd_in;
   int rc = run(h, MB_ABA, n, ld, q, qd, tau, fext, qdd, 0u, stream, opt);
   if (rc || !tau_out)
      return rc;
   const size_t row = (size_t)ld * sizeof(double);
   if (h->n_accel_source == 0)
   {
      if (tau_out != tau)
         MB_CUDA(h, cudaMemcpy2DAsync(tau_out, row, tau, row, (size_t)n * sizeof(double), (size_t)h->tree.nv, cudaMemcpyDeviceToDevice, stream));
      return MECANO_B200_OK;
   }
   if (tau_out == tau)
      return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "tau_out must not alias tau when joints are in ACCELERATION_SOURCE mode");
   rc = run(h, MB_RNEA, n, ld, q, qd, qdd, fext, tau_out, 0u, stream);
   if (rc) return rc;
   for (const auto &r : h->effort_dof_runs)
      MB_CUDA(h, cudaMemcpy2DAsync(tau_out + (size_t)r.first * ld, row, tau + (size_t)r.first * ld, row, (size_t)n * sizeof(double), (size_t)r.second,
                                   cudaMemcpyDeviceToDevice, stream));
   return MECANO_B200_OK;
}

// "rnea=512:32,aba=256:0" -> block size and TMEM stack slots of the specialised kernel of one algorithm
bool spec_cfg_from_env(int algo, mb::SpecOptions &opt)
{
   const char *e = getenv("MECANO_B200_SPEC_CFG");
   if (!e) return false;
   const char *key = algo == MB_RNEA ? "rnea=" : (algo == MB_ABA ? "aba=" : "crba=");
   const char *p = strstr(e, key);
   if (!p) return false;
   int b = 0, tm = 0, sync = 1;
   if (sscanf(p + strlen(key), "%d:%d:%d", &b, &tm, &sync) < 1 || b < 32 || b > 1024 || (b & 31)) return false;
   opt.block = b;
   opt.tm = tm;
   opt.sync_every = sync;
   return true;
}

int ensure_pipeline(mecano_b200_handle *h, size_t doubles_per_slot)
{
   for (int i = 0; i < h->n_slots; i++)
   {
      if (!h->streams[i]) MB_CUDA(h, cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking));
      if (!h->done[i]) MB_CUDA(h, cudaEventCreateWithFlags(&h->done[i], cudaEventDisableTiming));
   }
   if (doubles_per_slot > h->stage_doubles)
   {
      for (int i = 0; i < h->n_slots; i++)
      {
         if (h->stage[i]) cudaFree(h->stage[i]);
         h->stage[i] = nullptr;
      }
      h->stage_doubles = 0;
      for (int i = 0; i < h->n_slots; i++)
         MB_CUDA(h, cudaMalloc(&h->stage[i], doubles_per_slot * sizeof(double)));
      h->stage_doubles = doubles_per_slot;
   }
   return MECANO_B200_OK;
}

// rows x [s0, s0 + w) of a host matrix with leading dimension ld  <->  rows x w device matrix (ld = chunk)
cudaError_t copy_rows(double *dst, size_t dpitch, const double *src, size_t spitch, size_t w, size_t rows, cudaMemcpyKind kind, cudaStream_t s)
{
   return cudaMemcpy2DAsync(dst, dpitch * sizeof(double), src, spitch * sizeof(double), w * sizeof(double), rows, kind, s);
}

// ------------------------------------------------------------------------------------------------ host-pointer pipeline
// One host-pointer call on one handle ("lane"): pointers already offset to the lane's first state.  MB_STEP = the fused call
// of the three calculators (mecano_b200_step_host).
constexpr int MB_STEP = 100;
struct HostJob
{
   int algo = MB_RNEA;
   int64_t n = 0, ld = 0;
   int64_t variant_n = 0;                               // size of the whole call (all lanes)
   const double *q = nullptr, *qd = nullptr, *x = nullptr, *fext = nullptr, *x2 = nullptr; // x: qdd (RNEA) / tau (ABA); MB_STEP: qdd_in
   const double *tau_in = nullptr;                       // MB_STEP: efforts for forward dynamics
   double *out = nullptr;                                // tau (RNEA) / qdd (ABA) / mass matrix (CRBA); MB_STEP: tau_out
   double *body_acc = nullptr, *joint_wrench = nullptr, *tau_out = nullptr; // by-products (RNEA), efforts of pass four (ABA with source modes)
   double *qdd_out = nullptr, *M = nullptr;              // MB_STEP
   uint32_t flags = 0;                                   // RNEA flags / CRBA layout (MB_STEP: layout)
   // pipeline state
   size_t chunk = 0, in_rows = 0, out_rows = 0, m_rows = 0;
   int64_t s0 = 0;
   int slot = 0;
};

size_t mass_matrix_rows(const mecano_b200_handle *h, uint32_t layout)
{
   return (layout & MECANO_B200_CRBA_PACKED) ? h->tree.packed_row.size() : (size_t)h->tree.nv * (size_t)h->tree.nv;
}

// argument checks, chunk size, staging buffers.  Two slots per handle, each with its own stream: H2D of chunk k+1 overlaps the
// kernels and the D2H of chunk k.
int host_begin(mecano_b200_handle *h, HostJob &j)
{
   int rc = check_batch(h, j.n, j.ld);
   if (rc) return rc;
   if (j.n == 0) return MECANO_B200_OK;
   const size_t nq = h->tree.nq, nv = h->tree.nv, nb = h->tree.nb;
   if (j.algo == MB_STEP)
   {
      const bool rnea = j.x || j.out, aba = j.tau_in || j.qdd_out;
      if (!j.q || ((rnea || aba) && !j.qd) || (rnea && (!j.x || !j.out)) || (aba && (!j.tau_in || !j.qdd_out)) || (!rnea && !aba && !j.M))
         return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "NULL buffer");
      if (aba && h->n_accel_source > 0)
         return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "joints in ACCELERATION_SOURCE mode need their accelerations: call mecano_b200_aba_sources_host");
      // (state-major: as below, the kernel rewrites the structural zeros in the staging buffer every time)
      if (j.flags & MECANO_B200_CRBA_STATE_MAJOR)
         j.flags &= ~MECANO_B200_CRBA_ZEROS_PRESENT;
      j.m_rows = j.M ? mass_matrix_rows(h, j.flags) : 0;
      j.in_rows = nq + (rnea || aba ? nv : 0) + (rnea ? nv : 0) + (aba ? nv : 0) + (j.fext ? 6 * nb : 0);
      j.out_rows = (rnea ? nv : 0) + (aba ? nv : 0) + j.m_rows;
   }
   else
   {
      if (!j.q || !j.out || (j.algo != MB_CRBA && (!j.qd || !j.x)))
         return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "NULL buffer");
      // state-major: the whole per-state block is copied back from the staging buffer, which other calls on this handle reuse, so
      // the kernel must write the structural zeros there every time (the transfer saving of ZEROS_PRESENT is entry-major only)
      if (j.algo == MB_CRBA && (j.flags & MECANO_B200_CRBA_STATE_MAJOR))
         j.flags &= ~MECANO_B200_CRBA_ZEROS_PRESENT;
      j.m_rows = j.algo == MB_CRBA ? mass_matrix_rows(h, j.flags) : 0;
      j.in_rows = j.algo == MB_CRBA ? nq : nq + 2 * nv + (j.fext ? 6 * nb : 0) + (j.x2 ? nv : 0);
      j.out_rows = j.algo == MB_CRBA ? j.m_rows : nv + (j.body_acc ? 6 * nb : 0) + (j.joint_wrench ? 6 * nb : 0) + (j.tau_out ? nv : 0);
   }
   if ((j.flags & MECANO_B200_CRBA_PACKED) && (j.flags & (MECANO_B200_CRBA_STATE_MAJOR | MECANO_B200_CRBA_ZEROS_PRESENT)) && (j.algo == MB_CRBA || j.algo == MB_STEP))
      return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "the packed mass-matrix layout does not combine with STATE_MAJOR / ZEROS_PRESENT");
   // chunk: ~256 MB of rows per slot (MECANO_B200_HOST_CHUNK_MB; r06ze: 229 ms per dense 2^20-state step against 239 ms with 128 MB, 231 ms with
   // 512 MB), at least 4096 states, multiple of 256
   static const double chunk_mb = [] { const char *e = getenv("MECANO_B200_HOST_CHUNK_MB"); const double v = e ? atof(e) : 0.0; return v >= 1.0 ? v : 256.0; }();
   size_t chunk = (size_t)(chunk_mb * 1024 * 1024 / 8 / (double)(j.in_rows + j.out_rows));
   chunk = std::max<size_t>(4096, chunk & ~(size_t)255);
   chunk = std::min<size_t>(chunk, ((size_t)j.n + 255) & ~(size_t)255);
   j.chunk = chunk;
   j.s0 = 0;
   j.slot = 0;
   MB_ON_DEVICE(h);
   return ensure_pipeline(h, (j.in_rows + j.out_rows) * chunk);
}

// issues the next chunk of the job (H2D, kernels, D2H on the slot's stream); asynchronous with pinned host memory
int host_issue_chunk(mecano_b200_handle *h, HostJob &j)
{
   if (j.s0 >= j.n) return MECANO_B200_OK;
   MB_ON_DEVICE(h);
   const size_t nq = h->tree.nq, nv = h->tree.nv, nb = h->tree.nb, chunk = j.chunk, ld = (size_t)j.ld;
   const int64_t s0 = j.s0;
   const size_t w = (size_t)std::min<int64_t>((int64_t)chunk, j.n - s0);
   const int slot = j.slot;
   cudaStream_t st = h->streams[slot];
   j.s0 += (int64_t)chunk;
   j.slot = (slot + 1) % h->n_slots;
   // a slot is reused every n_slots chunks: stream order already serialises it
   double *p = h->stage[slot];
   auto take = [&](size_t rows) { double *r = p; p += rows * chunk; return r; };
   auto h2d = [&](double *dst, const double *src, size_t rows) { return copy_rows(dst, chunk, src + s0, ld, w, rows, cudaMemcpyHostToDevice, st); };
   auto d2h = [&](double *dst, const double *src, size_t rows) { return copy_rows(dst + s0, ld, src, chunk, w, rows, cudaMemcpyDeviceToHost, st); };
   int rc = MECANO_B200_OK;
   RunOpts opt;
   opt.ws_slot = 1 + slot;
   opt.variant_n = j.variant_n > 0 ? j.variant_n : j.n;
   if (j.algo == MB_STEP)
   {
      const bool rnea = j.x != nullptr, aba = j.tau_in != nullptr;
      double *dq = take(nq), *dqd = (rnea || aba) ? take(nv) : nullptr, *dqdd = rnea ? take(nv) : nullptr, *dtin = aba ? take(nv) : nullptr;
      double *df = j.fext ? take(6 * nb) : nullptr;
      double *dtau = rnea ? take(nv) : nullptr, *dqo = aba ? take(nv) : nullptr, *dM = j.M ? take(j.m_rows) : nullptr;
      MB_CUDA(h, h2d(dq, j.q, nq));
      if (dqd) MB_CUDA(h, h2d(dqd, j.qd, nv));
      if (dqdd) MB_CUDA(h, h2d(dqdd, j.x, nv));
      if (dtin) MB_CUDA(h, h2d(dtin, j.tau_in, nv));
      if (df) MB_CUDA(h, h2d(df, j.fext, 6 * nb));
      if (rnea)
      {
         rc = run(h, MB_RNEA, (int64_t)w, (int64_t)chunk, dq, dqd, dqdd, df, dtau, 0u, st, opt);
         if (rc) return rc;
         MB_CUDA(h, d2h(j.out, dtau, nv));
      }
      if (aba)
      {
         rc = run(h, MB_ABA, (int64_t)w, (int64_t)chunk, dq, dqd, dtin, df, dqo, 0u, st, opt);
         if (rc) return rc;
         MB_CUDA(h, d2h(j.qdd_out, dqo, nv));
      }
      if (j.M)
      {
         rc = run(h, MB_CRBA, (int64_t)w, (int64_t)chunk, dq, nullptr, nullptr, nullptr, dM, j.flags, st, opt);
         if (rc) return rc;
         if (j.flags & MECANO_B200_CRBA_STATE_MAJOR)
            MB_CUDA(h, cudaMemcpyAsync(j.M + (size_t)s0 * nv * nv, dM, w * nv * nv * sizeof(double), cudaMemcpyDeviceToHost, st));
         else if (j.flags & MECANO_B200_CRBA_ZEROS_PRESENT)
         {
            // entry-major: a structurally zero entry is a whole row of the host matrix, which already holds zeros (and which the
            // kernel did not write in the staging buffer)
            for (const auto &run : h->nonzero_runs)
               MB_CUDA(h, copy_rows(j.M + (size_t)run.first * ld + s0, ld, dM + (size_t)run.first * chunk, chunk, w, (size_t)run.second, cudaMemcpyDeviceToHost, st));
         }
         else
            MB_CUDA(h, d2h(j.M, dM, j.m_rows));
      }
      return MECANO_B200_OK;
   }
   const int algo = j.algo;
   const bool state_major = algo == MB_CRBA && (j.flags & MECANO_B200_CRBA_STATE_MAJOR);
   const bool sources = algo == MB_ABA && (j.x2 || j.tau_out); // mecano_b200_aba_sources_host
   double *dq = take(nq), *dqd = nullptr, *dx = nullptr, *df = nullptr, *dx2 = nullptr;
   MB_CUDA(h, h2d(dq, j.q, nq));
   if (algo != MB_CRBA)
   {
      dqd = take(nv);
      dx = take(nv);
      MB_CUDA(h, h2d(dqd, j.qd, nv));
      MB_CUDA(h, h2d(dx, j.x, nv));
      if (j.fext) { df = take(6 * nb); MB_CUDA(h, h2d(df, j.fext, 6 * nb)); }
      if (j.x2) { dx2 = take(nv); MB_CUDA(h, h2d(dx2, j.x2, nv)); }
   }
   double *dout = take(algo == MB_CRBA ? j.m_rows : nv);
   double *dacc = j.body_acc ? take(6 * nb) : nullptr;
   double *dwr = j.joint_wrench ? take(6 * nb) : nullptr;
   double *dtau = j.tau_out ? take(nv) : nullptr;
   if (sources)
      rc = run_aba_sources(h, (int64_t)w, (int64_t)chunk, dq, dqd, dx, dx2, df, dout, dtau, st, 1 + slot);
   else
   {
      opt.body_acc = dacc;
      opt.joint_wrench = dwr;
      rc = run(h, algo, (int64_t)w, (int64_t)chunk, dq, dqd, dx, df, dout, j.flags, st, opt);
   }
   if (rc) return rc;
   if (dtau) MB_CUDA(h, d2h(j.tau_out, dtau, nv));
   if (dacc) MB_CUDA(h, d2h(j.body_acc, dacc, 6 * nb));
   if (dwr) MB_CUDA(h, d2h(j.joint_wrench, dwr, 6 * nb));
   if (state_major)
      MB_CUDA(h, cudaMemcpyAsync(j.out + (size_t)s0 * nv * nv, dout, w * nv * nv * sizeof(double), cudaMemcpyDeviceToHost, st));
   else if (algo == MB_CRBA && (j.flags & MECANO_B200_CRBA_ZEROS_PRESENT))
   {
      // entry-major: a structurally zero entry is a whole row of the host matrix, which already holds zeros
      for (const auto &run : h->nonzero_runs)
         MB_CUDA(h, copy_rows(j.out + (size_t)run.first * ld + s0, ld, dout + (size_t)run.first * chunk, chunk, w, (size_t)run.second, cudaMemcpyDeviceToHost, st));
   }
   else
      MB_CUDA(h, d2h(j.out, dout, algo == MB_CRBA ? j.m_rows : nv));
   return MECANO_B200_OK;
}

int host_finish(mecano_b200_handle *h)
{
   MB_ON_DEVICE(h);
   for (int i = 0; i < h->n_slots; i++)
      if (h->streams[i]) MB_CUDA(h, cudaStreamSynchronize(h->streams[i]));
   return MECANO_B200_OK;
}

// Runs one job per lane: the chunks of all lanes are issued round-robin from the calling thread (everything is asynchronous
// with pinned host memory), then every lane is synchronised once.  lanes.size() == 1 is the single-device host call.
int run_host_lanes(const std::vector<mecano_b200_handle *> &lanes, std::vector<HostJob> &jobs, int *failed_lane)
{
   int rc = MECANO_B200_OK;
   size_t started = 0;
   *failed_lane = 0;
   // a listed handle is locked for the whole call (handles are distinct objects even when they share a device)
   for (; started < lanes.size(); started++)
   {
      lanes[started]->mu.lock();
      rc = host_begin(lanes[started], jobs[started]);
      if (rc) { *failed_lane = (int)started; started++; break; }
   }
   bool pending = rc == MECANO_B200_OK;
   while (pending && rc == MECANO_B200_OK)
   {
      pending = false;
      for (size_t i = 0; i < lanes.size() && rc == MECANO_B200_OK; i++)
      {
         if (jobs[i].s0 >= jobs[i].n) continue;
         rc = host_issue_chunk(lanes[i], jobs[i]);
         if (rc) *failed_lane = (int)i;
         pending = pending || jobs[i].s0 < jobs[i].n;
      }
   }
   // always drain what was issued, also after an error: the staging buffers and the caller's matrices are in flight
   for (size_t i = 0; i < started; i++)
   {
      if (jobs[i].n > 0 && lanes[i]->streams[0])
      {
         const int rf = host_finish(lanes[i]);
         if (rf && !rc) { rc = rf; *failed_lane = (int)i; }
      }
      lanes[i]->mu.unlock();
   }
   return rc;
}

int run_host(mecano_b200_handle *h, HostJob &job)
{
   if (!h) return MECANO_B200_ERR_INVALID_ARGUMENT;
   std::vector<mecano_b200_handle *> lanes(1, h);
   std::vector<HostJob> jobs(1, job);
   int failed = 0;
   return run_host_lanes(lanes, jobs, &failed);
}

int run_host(mecano_b200_handle *h, int algo, int64_t n, int64_t ld, const double *q, const double *qd, const double *x, const double *fext,
             double *out, uint32_t flags, double *body_acc = nullptr, double *joint_wrench = nullptr, const double *x2 = nullptr,
             double *tau_out = nullptr)
{
   HostJob j;
   j.algo = algo; j.n = n; j.ld = ld;
   j.q = q; j.qd = qd; j.x = x; j.fext = fext; j.x2 = x2;
   j.out = out; j.body_acc = body_acc; j.joint_wrench = joint_wrench; j.tau_out = tau_out;
   j.flags = flags;
   return run_host(h, j);
}
} // namespace

extern "C" {
#pragma GCC visibility push(default)

int mecano_b200_version(void) { return MECANO_B200_VERSION; }

int mecano_b200_device_count(void)
{
   int n = 0;
   if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
   return n;
}

int mecano_b200_create(const mecano_b200_tree_desc *desc, int device, mecano_b200_handle **out)
{
   if (!out) return fail(nullptr, MECANO_B200_ERR_INVALID_ARGUMENT, "out handle pointer is NULL");
   *out = nullptr;
   mecano_b200_handle *h = new mecano_b200_handle();
   std::string err;
   int rc = mb::flatten_tree(desc, h->tree, err);
   if (rc != MECANO_B200_OK)
   {
      delete h;
      return fail(nullptr, rc, err);
   }
   int ndev = 0;
   cudaError_t e = cudaGetDeviceCount(&ndev);
   if (e != cudaSuccess || ndev == 0)
   {
      delete h;
      return fail(nullptr, MECANO_B200_ERR_NO_DEVICE,
                  std::string("no CUDA device available (this engine has no CPU fallback)") + (e != cudaSuccess ? std::string(": ") + cudaGetErrorString(e) : ""));
   }
   if (device < 0 || device >= ndev)
   {
      delete h;
      return fail(nullptr, MECANO_B200_ERR_INVALID_ARGUMENT, "device index out of range");
   }
   h->device = device;
   if (const char *es = getenv("MECANO_B200_HOST_SLOTS"))
      h->n_slots = std::min(MB_MAX_SLOTS, std::max(1, atoi(es)));
   auto bail = [&](cudaError_t ce, const char *what) {
      std::string m = std::string(what) + ": " + cudaGetErrorString(ce);
      if (h->d_consts) cudaFree(h->d_consts);
      if (h->d_counters) cudaFree(h->d_counters);
      if (h->d_zero) cudaFree(h->d_zero);
      if (h->d_prog) cudaFree(h->d_prog);
      delete h;
      return fail(nullptr, (int)ce, m);
   };
   DeviceGuard device_guard_(device);
   if ((e = device_guard_.err) != cudaSuccess) return bail(e, "cudaSetDevice");
   const size_t bytes = h->tree.consts.size() * sizeof(double);
   if ((e = cudaMalloc(&h->d_consts, bytes)) != cudaSuccess) return bail(e, "cudaMalloc(consts)");
   if ((e = cudaMalloc(&h->d_counters, 2 * sizeof(unsigned) * MB_COUNTERS)) != cudaSuccess) return bail(e, "cudaMalloc(counters)");
   if ((e = cudaMemset(h->d_counters, 0, 2 * sizeof(unsigned) * MB_COUNTERS)) != cudaSuccess) return bail(e, "cudaMemset(counters)");
   if ((e = cudaMemcpy(h->d_consts, h->tree.consts.data(), bytes, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "cudaMemcpy(consts)");
   {
      std::vector<char> isz((size_t)h->tree.nv * h->tree.nv, 0);
      for (uint16_t e : h->tree.zero_entries)
         isz[e] = 1;
      for (int e = 0; e < (int)isz.size(); e++)
      {
         if (isz[(size_t)e])
            continue;
         if (!h->nonzero_runs.empty() && h->nonzero_runs.back().first + h->nonzero_runs.back().second == e)
            h->nonzero_runs.back().second++;
         else
            h->nonzero_runs.emplace_back(e, 1);
      }
   }
   if (!h->tree.zero_entries.empty())
   {
      const size_t zb = h->tree.zero_entries.size() * sizeof(uint16_t);
      if ((e = cudaMalloc(&h->d_zero, zb)) != cudaSuccess) return bail(e, "cudaMalloc(zero entries)");
      if ((e = cudaMemcpy(h->d_zero, h->tree.zero_entries.data(), zb, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "cudaMemcpy(zero entries)");
   }
   for (int algo = 0; algo < MB_NUM_ALGOS; algo++)
   {
      bool fits = false;
      if ((e = mb::plan_thread_kernel(algo, h->tree.prog[algo], false, h->plan[algo], &fits)) != cudaSuccess) return bail(e, "kernel planning");
      if (!fits && algo == MB_CORIOLIS)
      {
         h->plan[algo].block = 0; // the by-product kernel alone does not fit: mecano_b200_coriolis reports it, everything else works
         continue;
      }
      if (!fits)
      {
         cudaFree(h->d_consts);
         if (h->d_counters) cudaFree(h->d_counters);
         if (h->d_zero) cudaFree(h->d_zero);
         delete h;
         return fail(nullptr, MECANO_B200_ERR_TOO_LARGE, "tree exceeds the compiled per-state work areas (branch nesting / depth too large)");
      }
   }
   {
      const MbProgram &P = h->tree.prog[MB_RNEA];
      std::vector<int> nchild(P.nb, 0);
      for (int i = 0; i < P.nb; i++)
      {
         if (P.body[i].parent >= 0) nchild[P.body[i].parent]++;
         h->max_ndof = std::max(h->max_ndof, P.body[i].ndof);
         h->has_3dof = h->has_3dof || P.body[i].sub != MB_SUB_SIX;
      }
      for (int i = 0; i < P.nb; i++) h->max_children = std::max(h->max_children, nchild[i]);
      cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
      h->warp_ok = mb::warp_variant_supports(P);
      if (h->warp_ok)
      {
         if ((e = cudaMalloc(&h->d_prog, 3 * sizeof(MbProgram))) != cudaSuccess) return bail(e, "cudaMalloc(programs)");
         if ((e = cudaMemcpy(h->d_prog, h->tree.prog, 3 * sizeof(MbProgram), cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "cudaMemcpy(programs)");
      }
      // crossover batch sizes: linear in the body count through the measured crossovers of A7 (7 bodies,
      // profiles/r01h_batch_sweep_graph.jsonl) and H37 (32 bodies, profiles/r01y_batch_sweep.jsonl: the thread-per-state walk
      // of one state now takes 49 / 96 / 98 us, the warp kernels cross it at 4.8 k / 3.9 k / 3.4 k states).
      // MECANO_B200_WARP_BELOW overrides all three
      h->warp_below[MB_RNEA] = 2048 + 80 * P.nb; h->warp_below[MB_ABA] = 1536 + 64 * P.nb; h->warp_below[MB_CRBA] = 2048 + 40 * P.nb;
      if (P.nb > 32)
      {
         // team-per-state kernels (team_kernels.cu), trees of 51 / 75 / 101 bodies: crossovers at 2.5-2.7 k (RNEA), 1.5-2.0 k (ABA)
         // and 3.3-4.5 k states (CRBA), profiles/r05b_config5_1gpu.jsonl
         h->warp_below[MB_RNEA] = 2560; h->warp_below[MB_ABA] = 1792; h->warp_below[MB_CRBA] = 3584;
      }
      if (const char *e = getenv("MECANO_B200_WARP_BELOW"))
         h->warp_below[0] = h->warp_below[1] = h->warp_below[2] = atoll(e);
   }
   *out = h;
   return MECANO_B200_OK;
}

void mecano_b200_destroy(mecano_b200_handle *h)
{
   if (!h) return;
   DeviceGuard device_guard_(h->device);
   for (int i = 0; i < MB_MAX_SLOTS; i++)
   {
      if (h->stage[i]) cudaFree(h->stage[i]);
      if (h->done[i]) cudaEventDestroy(h->done[i]);
      if (h->streams[i]) cudaStreamDestroy(h->streams[i]);
   }
   if (h->d_zero_row) cudaFree(h->d_zero_row);
   if (h->d_scratch) cudaFree(h->d_scratch);
   if (h->d_consts) cudaFree(h->d_consts);
   if (h->d_counters) cudaFree(h->d_counters);
   if (h->d_zero) cudaFree(h->d_zero);
   if (h->d_prog) cudaFree(h->d_prog);
   for (int i = 0; i < 1 + MB_MAX_SLOTS; i++)
      if (h->d_ws[i]) cudaFree(h->d_ws[i]);
   for (int i = 0; i < 3; i++)
      mb::spec_unload(h->spec[i]);
   delete h;
}

const char *mecano_b200_last_error(const mecano_b200_handle *h)
{
   if (h) return h->error.c_str();
   return g_create_error.c_str();
}

int mecano_b200_set_gravity(mecano_b200_handle *h, double gx, double gy, double gz)
{
   if (!h) return MECANO_B200_ERR_INVALID_ARGUMENT;
   h->gravity[0] = gx; h->gravity[1] = gy; h->gravity[2] = gz;
   return MECANO_B200_OK;
}

int mecano_b200_set_precision(mecano_b200_handle *h, int precision)
{
   if (!h || (precision != MECANO_B200_PRECISION_FP64 && precision != MECANO_B200_PRECISION_FP32)) return MECANO_B200_ERR_INVALID_ARGUMENT;
   h->fp32 = precision == MECANO_B200_PRECISION_FP32;
   return MECANO_B200_OK;
}

int mecano_b200_set_grid_limit(mecano_b200_handle *h, int algo, int max_blocks)
{
   if (!h || algo < 0 || algo >= MB_NUM_ALGOS || max_blocks < 0) return MECANO_B200_ERR_INVALID_ARGUMENT;
   h->grid_limit[algo] = max_blocks;
   return MECANO_B200_OK;
}

int mecano_b200_set_variant(mecano_b200_handle *h, int variant)
{
   if (!h) return MECANO_B200_ERR_INVALID_ARGUMENT;
   if (variant != MECANO_B200_VARIANT_AUTO && variant != MECANO_B200_VARIANT_THREAD && variant != MECANO_B200_VARIANT_WARP)
      return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "unknown variant");
   if (variant == MECANO_B200_VARIANT_WARP && !h->warp_ok)
      return fail(h, MECANO_B200_ERR_UNSUPPORTED_TOPOLOGY, "the warp-per-state variant (one lane per body) handles trees of one-DoF and SixDoF joints");
   h->variant = variant;
   return MECANO_B200_OK;
}

int mecano_b200_specialize(mecano_b200_handle *h, uint32_t algo_mask)
{
   if (!h) return MECANO_B200_ERR_INVALID_ARGUMENT;
   std::lock_guard<std::mutex> lk(h->mu);
   MB_ON_DEVICE(h);
   for (int algo = 0; algo < 3; algo++)
   {
      if (!(algo_mask & (1u << algo)) || h->spec[algo].ready())
         continue;
      if (algo == MB_CRBA)
         continue; // bound by the HBM write of the mass matrix: the generic kernel is already at the roof
      if (h->has_3dof)
         continue; // spherical / planar joints: the generic kernels serve them (two-slot pass-three records, multidof_aba.cuh)
      const MbProgram &P = h->tree.prog[algo];
      // Unrolled code is fetched once per tile of states; beyond ~100 KB it no longer fits the instruction caches and the
      // block becomes fetch-bound (measured: 32-body humanoid RNEA 172 KB -> 0.78x, 15-body tree 81 KB -> 1.42x of the generic
      // kernel; profiles/r01g_spec.jsonl).  Estimated size: RNEA ~5.4 KB, ABA ~12 KB of SASS per body.
      const double est_kb = h->tree.nb * (algo == MB_RNEA ? 5.4 : 12.0);
      if (est_kb > 110.0 && !(algo_mask & MECANO_B200_SPECIALIZE_FORCE))
         continue;
      mb::SpecOptions opt;
      // defaults from the launch-configuration sweep (profiles/): RNEA 16 warps with a TMEM stack, ABA 8 warps in shared memory
      if (algo == MB_RNEA) { opt.block = 512; opt.tm = 32; }
      else { opt.block = 256; opt.tm = 0; }
      spec_cfg_from_env(algo, opt);
      if (!mb_tm_fits(algo, P, opt.tm, opt.block))
         opt.tm = 0; // the wide stack area of a deep tree does not fit the TMEM columns of one warp: shared memory only
      if (opt.tm * 4 > (512 / ((opt.block + 127) / 128) & ~3))
         return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "MECANO_B200_SPEC_CFG: TMEM slots exceed the columns of one warp");
      std::string err;
      int rc = MECANO_B200_ERR_TOO_LARGE;
      // shrink the block until the shared-memory part of the stack fits
      for (int b = opt.block; b >= 64 && rc == MECANO_B200_ERR_TOO_LARGE; b >>= 1)
      {
         mb::SpecOptions o2 = opt;
         o2.block = b;
         o2.tm = opt.tm > 0 ? std::min(opt.tm * (opt.block / b), 128) : 0;
         if (!mb_tm_fits(algo, P, o2.tm, o2.block))
            o2.tm = 0;
         int max_optin = 0;
         cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
         if (mb::spec_smem_bytes(algo, P, o2.block, o2.tm) + 1024 > (size_t)max_optin)
            continue;
         rc = mb::spec_build(algo, h->tree, o2, h->spec[algo], err);
      }
      if (rc != MECANO_B200_OK)
         return fail(h, rc, "specialize: " + err);
   }
   return MECANO_B200_OK;
}

int mecano_b200_n_dofs(const mecano_b200_handle *h) { return h ? h->tree.nv : -1; }
int mecano_b200_n_cfg(const mecano_b200_handle *h) { return h ? h->tree.nq : -1; }
int mecano_b200_n_bodies(const mecano_b200_handle *h) { return h ? h->tree.nb : -1; }

int mecano_b200_rnea(mecano_b200_handle *h, int64_t n, int64_t ld, const double *q, const double *qd, const double *qdd, const double *fext,
                     double *tau, uint32_t flags, void *stream)
{
   int rc = check_batch(h, n, ld);
   if (rc) return rc;
   if (n == 0) return MECANO_B200_OK; /* empty batch: nothing to do, pointers may be NULL */
   if (!q || !qd || !qdd || !tau) return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "NULL buffer");
   MB_ON_DEVICE(h);
   return run(h, MB_RNEA, n, ld, q, qd, qdd, fext, tau, flags, (cudaStream_t)stream);
}

int mecano_b200_rnea_full(mecano_b200_handle *h, int64_t n, int64_t ld, const double *q, const double *qd, const double *qdd, const double *fext,
                          double *tau, double *body_acc, double *joint_wrench, uint32_t flags, void *stream)
{
   int rc = check_batch(h, n, ld);
   if (rc) return rc;
   if (n == 0) return MECANO_B200_OK;
   if (!q || !qd || !qdd || !tau) return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "NULL buffer");
   MB_ON_DEVICE(h);
   RunOpts opt;
   opt.body_acc = body_acc;
   opt.joint_wrench = joint_wrench;
   return run(h, MB_RNEA, n, ld, q, qd, qdd, fext, tau, flags, (cudaStream_t)stream, opt);
}

int mecano_b200_aba(mecano_b200_handle *h, int64_t n, int64_t ld, const double *q, const double *qd, const double *tau, const double *fext,
                    double *qdd, uint32_t flags, void *stream)
{
   int rc = check_batch(h, n, ld);
   if (rc) return rc;
   if (n == 0) return MECANO_B200_OK;
   if (!q || !qd || !tau || !qdd) return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "NULL buffer");
   MB_ON_DEVICE(h);
   return run(h, MB_ABA, n, ld, q, qd, tau, fext, qdd, flags, (cudaStream_t)stream);
}

int mecano_b200_set_joint_source_modes(mecano_b200_handle *h, const int32_t *accel_source)
{
   if (!h) return MECANO_B200_ERR_INVALID_ARGUMENT;
   std::lock_guard<std::mutex> lk(h->mu);
   h->n_accel_source = mb::apply_source_modes(h->tree, accel_source, h->effort_dof_runs);
   return MECANO_B200_OK;
}

int mecano_b200_aba_sources(mecano_b200_handle *h, int64_t n, int64_t ld, const double *q, const double *qd, const double *tau, const double *qdd_in,
                            const double *fext, double *qdd, double *tau_out, void *stream)
{
   int rc = check_batch(h, n, ld);
   if (rc) return rc;
   if (n == 0) return MECANO_B200_OK;
   if (!q || !qd || !tau || !qdd) return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "NULL buffer");
   MB_ON_DEVICE(h);
   return run_aba_sources(h, n, ld, q, qd, tau, qdd_in, fext, qdd, tau_out, (cudaStream_t)stream, 0);
}

int mecano_b200_aba_sources_host(mecano_b200_handle *h, int64_t n, int64_t ld, const double *q, const double *qd, const double *tau,
                                 const double *qdd_in, const double *fext, double *qdd, double *tau_out)
{
   if (h && h->n_accel_source > 0 && !qdd_in)
      return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "NULL buffer (qdd_in) with joints in ACCELERATION_SOURCE mode");
   return run_host(h, MB_ABA, n, ld, q, qd, tau, fext, qdd, 0u, nullptr, nullptr, qdd_in, tau_out);
}

int mecano_b200_crba_centroidal(mecano_b200_handle *h, int64_t n, int64_t ld, const double *q, double *M, double *cmm, double *com, int frame,
                                void *stream)
{
   int rc = check_batch(h, n, ld);
   if (rc) return rc;
   if (n == 0) return MECANO_B200_OK;
   if (!q || !M || !cmm || !com) return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "NULL buffer");
   if (frame != MECANO_B200_FRAME_WORLD && frame != MECANO_B200_FRAME_CENTER_OF_MASS)
      return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "unknown centroidal momentum frame");
   MB_ON_DEVICE(h);
   cudaStream_t st = (cudaStream_t)stream;
   // the kernel sums (mass * CoM, mass) over the children of the root body into the com rows
   MB_CUDA(h, cudaMemset2DAsync(com, (size_t)ld * sizeof(double), 0, (size_t)n * sizeof(double), 4, st));
   RunOpts opt;
   opt.cmm = cmm;
   opt.com = com;
   rc = run(h, MB_CRBA, n, ld, q, nullptr, nullptr, nullptr, M, MECANO_B200_CRBA_ENTRY_MAJOR, st, opt);
   if (rc) return rc;
   mb::CentroidalArgs ca;
   ca.cols = cmm; ca.com = com; ca.n = n; ca.ld = ld; ca.ncols = h->tree.nv;
   ca.normalize_com = 1;
   ca.shift = frame == MECANO_B200_FRAME_CENTER_OF_MASS;
   MB_CUDA(h, mb::launch_centroidal_finish(ca, st));
   return MECANO_B200_OK;
}

// The centre of mass alone (CenterOfMassCalculator.getCenterOfMass() / getTotalMass(), CenterOfMassCalculator.java:70-124; what
// CenterOfMassReferenceFrame.updateTransformToParent(), CenterOfMassReferenceFrame.java:45-50, asks for): the by-product CRBA kernel
// launched without a matrix keeps the composite-inertia recursion and drops the unit momenta, their ancestor walks and every
// matrix store (crba.cuh: com_only) -- the same arithmetic, hence the same bits, as the com rows of mecano_b200_crba_centroidal.
int mecano_b200_center_of_mass(mecano_b200_handle *h, int64_t n, int64_t ld, const double *q, double *com, void *stream)
{
   int rc = check_batch(h, n, ld);
   if (rc) return rc;
   if (n == 0) return MECANO_B200_OK;
   if (!q || !com) return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "NULL buffer");
   MB_ON_DEVICE(h);
   cudaStream_t st = (cudaStream_t)stream;
   MB_CUDA(h, cudaMemset2DAsync(com, (size_t)ld * sizeof(double), 0, (size_t)n * sizeof(double), 4, st));
   RunOpts opt;
   opt.com = com;
   rc = run(h, MB_CRBA, n, ld, q, nullptr, nullptr, nullptr, nullptr, MECANO_B200_CRBA_ENTRY_MAJOR, st, opt);
   if (rc) return rc;
   mb::CentroidalArgs ca;
   ca.cols = nullptr; ca.com = com; ca.n = n; ca.ld = ld; ca.ncols = 0;
   ca.normalize_com = 1;
   ca.shift = 0;
   MB_CUDA(h, mb::launch_centroidal_finish(ca, st));
   return MECANO_B200_OK;
}

int mecano_b200_centroidal_convective_term(mecano_b200_handle *h, int64_t n, int64_t ld, const double *q, const double *qd, const double *com,
                                           double *out, int frame, void *stream)
{
   int rc = check_batch(h, n, ld);
   if (rc) return rc;
   if (n == 0) return MECANO_B200_OK;
   if (!q || !qd || !out) return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "NULL buffer");
   if (frame != MECANO_B200_FRAME_WORLD && frame != MECANO_B200_FRAME_CENTER_OF_MASS)
      return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "unknown centroidal momentum frame");
   if (frame == MECANO_B200_FRAME_CENTER_OF_MASS && !com)
      return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "the centre-of-mass frame needs the com rows of mecano_b200_crba_centroidal");
   MB_ON_DEVICE(h);
   cudaStream_t st = (cudaStream_t)stream;
   const size_t need = (size_t)h->tree.nv * (size_t)ld;
   if (h->scratch_doubles < need)
   {
      if (h->d_scratch)
      {
         MB_CUDA(h, cudaDeviceSynchronize());
         MB_CUDA(h, cudaFree(h->d_scratch));
         h->d_scratch = nullptr;
         h->scratch_doubles = 0;
      }
      MB_CUDA(h, cudaMalloc(&h->d_scratch, need * sizeof(double)));
      h->scratch_doubles = need;
   }
   // :811-839: the net wrench of every body under zero joint accelerations and no gravity, summed in the centroidal frame = the
   // wrench at the root of inverse dynamics run that way
   MB_CUDA(h, cudaMemset2DAsync(out, (size_t)ld * sizeof(double), 0, (size_t)n * sizeof(double), 6, st));
   RunOpts opt;
   opt.root_wrench = out;
   opt.zero_gravity = true;
   rc = run(h, MB_RNEA, n, ld, q, qd, qd, nullptr, h->d_scratch, MECANO_B200_RNEA_NO_ACCELERATIONS, st, opt);
   if (rc) return rc;
   if (frame == MECANO_B200_FRAME_CENTER_OF_MASS)
   {
      mb::CentroidalArgs ca;
      ca.cols = out; ca.com = const_cast<double *>(com); ca.n = n; ca.ld = ld; ca.ncols = 1;
      ca.normalize_com = 0;
      ca.shift = 1;
      MB_CUDA(h, mb::launch_centroidal_finish(ca, st));
   }
   return MECANO_B200_OK;
}

int mecano_b200_coriolis(mecano_b200_handle *h, int64_t n, int64_t ld, const double *q, const double *qd, double *M, double *C, void *stream)
{
   int rc = check_batch(h, n, ld);
   if (rc) return rc;
   if (n == 0) return MECANO_B200_OK;
   if (!q || !qd || !M || !C) return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "NULL buffer");
   if (h->plan[MB_CORIOLIS].block == 0)
      return fail(h, MECANO_B200_ERR_TOO_LARGE, "tree exceeds the per-state work areas of the Coriolis-matrix kernel (branch nesting / depth too large)");
   MB_ON_DEVICE(h);
   RunOpts opt;
   opt.cor = C;
   return run(h, MB_CORIOLIS, n, ld, q, qd, qd, nullptr, M, 0u, (cudaStream_t)stream, opt);
}

int mecano_b200_coriolis_host(mecano_b200_handle *h, int64_t n, int64_t ld, const double *q, const double *qd, double *M, double *C)
{
   int rc = check_batch(h, n, ld);
   if (rc) return rc;
   if (n == 0) return MECANO_B200_OK;
   if (!q || !qd || !M || !C) return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "NULL buffer");
   std::lock_guard<std::mutex> lk(h->mu);
   MB_ON_DEVICE(h);
   const size_t nq = h->tree.nq, nv = h->tree.nv, rows = nq + nv + 2 * nv * nv;
   const size_t chunk = (size_t)std::min<int64_t>(n, 16384);
   double *d = nullptr;
   MB_CUDA(h, cudaMalloc(&d, rows * chunk * sizeof(double)));
   double *dq = d, *dqd = dq + nq * chunk, *dM = dqd + nv * chunk, *dC = dM + nv * nv * chunk;
   cudaError_t e = cudaSuccess;
   for (int64_t s0 = 0; s0 < n && rc == 0 && e == cudaSuccess; s0 += (int64_t)chunk)
   {
      const size_t w = (size_t)std::min<int64_t>((int64_t)chunk, n - s0);
      e = copy_rows(dq, chunk, q + s0, (size_t)ld, w, nq, cudaMemcpyHostToDevice, nullptr);
      if (e == cudaSuccess) e = copy_rows(dqd, chunk, qd + s0, (size_t)ld, w, nv, cudaMemcpyHostToDevice, nullptr);
      if (e != cudaSuccess) break;
      rc = mecano_b200_coriolis(h, (int64_t)w, (int64_t)chunk, dq, dqd, dM, dC, nullptr);
      if (rc) break;
      e = copy_rows(M + s0, (size_t)ld, dM, chunk, w, nv * nv, cudaMemcpyDeviceToHost, nullptr);
      if (e == cudaSuccess) e = copy_rows(C + s0, (size_t)ld, dC, chunk, w, nv * nv, cudaMemcpyDeviceToHost, nullptr);
      if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
   }
   cudaFree(d);
   if (e != cudaSuccess) return cuda_fail(h, e, "mecano_b200_coriolis_host");
   return rc;
}

// host-pointer variants of the two centroidal calls: plain staging through temporary device buffers, chunk by chunk
int mecano_b200_crba_centroidal_host(mecano_b200_handle *h, int64_t n, int64_t ld, const double *q, double *M, double *cmm, double *com, int frame)
{
   int rc = check_batch(h, n, ld);
   if (rc) return rc;
   if (n == 0) return MECANO_B200_OK;
   if (!q || !M || !cmm || !com) return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "NULL buffer");
   std::lock_guard<std::mutex> lk(h->mu);
   MB_ON_DEVICE(h);
   const size_t nq = h->tree.nq, nv = h->tree.nv, rows = nq + nv * nv + 6 * nv + 4;
   const size_t chunk = (size_t)std::min<int64_t>(n, 16384);
   double *d = nullptr;
   MB_CUDA(h, cudaMalloc(&d, rows * chunk * sizeof(double)));
   double *dq = d, *dM = dq + nq * chunk, *dA = dM + nv * nv * chunk, *dc = dA + 6 * nv * chunk;
   cudaError_t e = cudaSuccess;
   for (int64_t s0 = 0; s0 < n && rc == 0 && e == cudaSuccess; s0 += (int64_t)chunk)
   {
      const size_t w = (size_t)std::min<int64_t>((int64_t)chunk, n - s0);
      e = copy_rows(dq, chunk, q + s0, (size_t)ld, w, nq, cudaMemcpyHostToDevice, nullptr);
      if (e != cudaSuccess) break;
      rc = mecano_b200_crba_centroidal(h, (int64_t)w, (int64_t)chunk, dq, dM, dA, dc, frame, nullptr);
      if (rc) break;
      e = copy_rows(M + s0, (size_t)ld, dM, chunk, w, nv * nv, cudaMemcpyDeviceToHost, nullptr);
      if (e == cudaSuccess) e = copy_rows(cmm + s0, (size_t)ld, dA, chunk, w, 6 * nv, cudaMemcpyDeviceToHost, nullptr);
      if (e == cudaSuccess) e = copy_rows(com + s0, (size_t)ld, dc, chunk, w, 4, cudaMemcpyDeviceToHost, nullptr);
      if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
   }
   cudaFree(d);
   if (e != cudaSuccess) return cuda_fail(h, e, "mecano_b200_crba_centroidal_host");
   return rc;
}

int mecano_b200_centroidal_convective_term_host(mecano_b200_handle *h, int64_t n, int64_t ld, const double *q, const double *qd, const double *com,
                                                double *out, int frame)
{
   int rc = check_batch(h, n, ld);
   if (rc) return rc;
   if (n == 0) return MECANO_B200_OK;
   if (!q || !qd || !out) return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "NULL buffer");
   std::lock_guard<std::mutex> lk(h->mu);
   MB_ON_DEVICE(h);
   const size_t nq = h->tree.nq, nv = h->tree.nv, rows = nq + nv + 4 + 6;
   const size_t chunk = (size_t)std::min<int64_t>(n, 262144);
   double *d = nullptr;
   MB_CUDA(h, cudaMalloc(&d, rows * chunk * sizeof(double)));
   double *dq = d, *dqd = dq + nq * chunk, *dc = dqd + nv * chunk, *dout = dc + 4 * chunk;
   cudaError_t e = cudaSuccess;
   for (int64_t s0 = 0; s0 < n && rc == 0 && e == cudaSuccess; s0 += (int64_t)chunk)
   {
      const size_t w = (size_t)std::min<int64_t>((int64_t)chunk, n - s0);
      e = copy_rows(dq, chunk, q + s0, (size_t)ld, w, nq, cudaMemcpyHostToDevice, nullptr);
      if (e == cudaSuccess) e = copy_rows(dqd, chunk, qd + s0, (size_t)ld, w, nv, cudaMemcpyHostToDevice, nullptr);
      if (e == cudaSuccess && com) e = copy_rows(dc, chunk, com + s0, (size_t)ld, w, 4, cudaMemcpyHostToDevice, nullptr);
      if (e != cudaSuccess) break;
      rc = mecano_b200_centroidal_convective_term(h, (int64_t)w, (int64_t)chunk, dq, dqd, com ? dc : nullptr, dout, frame, nullptr);
      if (rc) break;
      e = copy_rows(out + s0, (size_t)ld, dout, chunk, w, 6, cudaMemcpyDeviceToHost, nullptr);
      if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
   }
   cudaFree(d);
   if (e != cudaSuccess) return cuda_fail(h, e, "mecano_b200_centroidal_convective_term_host");
   return rc;
}

int mecano_b200_center_of_mass_host(mecano_b200_handle *h, int64_t n, int64_t ld, const double *q, double *com)
{
   int rc = check_batch(h, n, ld);
   if (rc) return rc;
   if (n == 0) return MECANO_B200_OK;
   if (!q || !com) return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "NULL buffer");
   std::lock_guard<std::mutex> lk(h->mu);
   MB_ON_DEVICE(h);
   const size_t nq = h->tree.nq, rows = nq + 4;
   const size_t chunk = (size_t)std::min<int64_t>(n, 262144);
   double *d = nullptr;
   MB_CUDA(h, cudaMalloc(&d, rows * chunk * sizeof(double)));
   double *dq = d, *dc = dq + nq * chunk;
   cudaError_t e = cudaSuccess;
   for (int64_t s0 = 0; s0 < n && rc == 0 && e == cudaSuccess; s0 += (int64_t)chunk)
   {
      const size_t w = (size_t)std::min<int64_t>((int64_t)chunk, n - s0);
      e = copy_rows(dq, chunk, q + s0, (size_t)ld, w, nq, cudaMemcpyHostToDevice, nullptr);
      if (e != cudaSuccess) break;
      rc = mecano_b200_center_of_mass(h, (int64_t)w, (int64_t)chunk, dq, dc, nullptr);
      if (rc) break;
      e = copy_rows(com + s0, (size_t)ld, dc, chunk, w, 4, cudaMemcpyDeviceToHost, nullptr);
      if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
   }
   cudaFree(d);
   if (e != cudaSuccess) return cuda_fail(h, e, "mecano_b200_center_of_mass_host");
   return rc;
}

int mecano_b200_crba(mecano_b200_handle *h, int64_t n, int64_t ld, const double *q, double *M, uint32_t layout, void *stream)
{
   int rc = check_batch(h, n, ld);
   if (rc) return rc;
   if (n == 0) return MECANO_B200_OK;
   if (!q || !M) return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "NULL buffer");
   MB_ON_DEVICE(h);
   return run(h, MB_CRBA, n, ld, q, nullptr, nullptr, nullptr, M, layout, (cudaStream_t)stream);
}

int mecano_b200_rnea_host(mecano_b200_handle *h, int64_t n, int64_t ld, const double *q, const double *qd, const double *qdd,
                          const double *fext, double *tau, uint32_t flags)
{
   return run_host(h, MB_RNEA, n, ld, q, qd, qdd, fext, tau, flags);
}

int mecano_b200_rnea_full_host(mecano_b200_handle *h, int64_t n, int64_t ld, const double *q, const double *qd, const double *qdd,
                               const double *fext, double *tau, double *body_acc, double *joint_wrench, uint32_t flags)
{
   return run_host(h, MB_RNEA, n, ld, q, qd, qdd, fext, tau, flags, body_acc, joint_wrench);
}

int mecano_b200_aba_host(mecano_b200_handle *h, int64_t n, int64_t ld, const double *q, const double *qd, const double *tau,
                         const double *fext, double *qdd, uint32_t flags)
{
   return run_host(h, MB_ABA, n, ld, q, qd, tau, fext, qdd, flags);
}

int mecano_b200_integrate(mecano_b200_handle *h, int64_t n, int64_t ld, double dt, double *q, double *qd, double *qdd, void *stream)
{
   int rc = check_batch(h, n, ld);
   if (rc) return rc;
   if (n == 0) return MECANO_B200_OK;
   if (!q || !qd || !qdd) return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "NULL buffer");
   if (!(dt == dt)) return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "dt is NaN");
   MB_ON_DEVICE(h);
   mb::IntegrateJoints J;
   const MbProgram &P = h->tree.prog[MB_RNEA];
   J.nb = P.nb;
   for (int i = 0; i < P.nb; i++)
   {
      J.cfg[i] = (uint16_t)P.body[i].cfg_off;
      J.dof[i] = (uint16_t)P.body[i].dof_off;
      J.type[i] = (uint8_t)P.body[i].jtype;
      J.sub[i] = (uint8_t)P.body[i].sub;
   }
   mb::IntegrateArgs a;
   a.q = q; a.qd = qd; a.qdd = qdd;
   a.n = n; a.ld = ld;
   a.dt = dt;
   MB_CUDA(h, mb::launch_integrate_kernel(J, a, h->sm_count, (cudaStream_t)stream));
   return MECANO_B200_OK;
}

int mecano_b200_integrate_host(mecano_b200_handle *h, int64_t n, int64_t ld, double dt, double *q, double *qd, double *qdd)
{
   int rc = check_batch(h, n, ld);
   if (rc) return rc;
   if (n == 0) return MECANO_B200_OK;
   if (!q || !qd || !qdd) return fail(h, MECANO_B200_ERR_INVALID_ARGUMENT, "NULL buffer");
   std::lock_guard<std::mutex> lk(h->mu);
   MB_ON_DEVICE(h);
   const size_t nq = h->tree.nq, nv = h->tree.nv, rows = nq + 2 * nv;
   size_t chunk = (size_t)(64.0 * 1024 * 1024 / 8 / (double)rows);
   chunk = std::max<size_t>(4096, chunk & ~(size_t)255);
   chunk = std::min<size_t>(chunk, ((size_t)n + 255) & ~(size_t)255);
   rc = ensure_pipeline(h, rows * chunk);
   if (rc) return rc;
   int slot = 0;
   for (int64_t s0 = 0; s0 < n; s0 += (int64_t)chunk, slot = (slot + 1) % h->n_slots)
   {
      const size_t w = (size_t)std::min<int64_t>((int64_t)chunk, n - s0);
      cudaStream_t st = h->streams[slot];
      double *dq = h->stage[slot], *dqd = dq + nq * chunk, *dx = dqd + nv * chunk;
      MB_CUDA(h, copy_rows(dq, chunk, q + s0, (size_t)ld, w, nq, cudaMemcpyHostToDevice, st));
      MB_CUDA(h, copy_rows(dqd, chunk, qd + s0, (size_t)ld, w, nv, cudaMemcpyHostToDevice, st));
      MB_CUDA(h, copy_rows(dx, chunk, qdd + s0, (size_t)ld, w, nv, cudaMemcpyHostToDevice, st));
      rc = mecano_b200_integrate(h, (int64_t)w, (int64_t)chunk, dt, dq, dqd, dx, st);
      if (rc) return rc;
      MB_CUDA(h, copy_rows(q + s0, (size_t)ld, dq, chunk, w, nq, cudaMemcpyDeviceToHost, st));
      MB_CUDA(h, copy_rows(qd + s0, (size_t)ld, dqd, chunk, w, nv, cudaMemcpyDeviceToHost, st));
      MB_CUDA(h, copy_rows(qdd + s0, (size_t)ld, dx, chunk, w, nv, cudaMemcpyDeviceToHost, st));
   }
   return host_finish(h);
}

int mecano_b200_crba_host(mecano_b200_handle *h, int64_t n, int64_t ld, const double *q, double *M, uint32_t layout)
{
   return run_host(h, MB_CRBA, n, ld, q, nullptr, nullptr, nullptr, M, layout);
}

int mecano_b200_crba_packed_size(const mecano_b200_handle *h) { return h ? (int)h->tree.packed_row.size() : -1; }

int mecano_b200_crba_packed_index(const mecano_b200_handle *h, int32_t *row, int32_t *col)
{
   if (!h) return MECANO_B200_ERR_INVALID_ARGUMENT;
   if (row) std::copy(h->tree.packed_row.begin(), h->tree.packed_row.end(), row);
   if (col) std::copy(h->tree.packed_col.begin(), h->tree.packed_col.end(), col);
   return MECANO_B200_OK;
}

int mecano_b200_step_host(mecano_b200_handle *h, int64_t n, int64_t ld, const double *q, const double *qd, const double *qdd_in, const double *tau_in,
                          const double *fext, double *tau_out, double *qdd_out, double *M, uint32_t layout)
{
   HostJob j;
   j.algo = MB_STEP; j.n = n; j.ld = ld;
   j.q = q; j.qd = qd; j.x = qdd_in; j.tau_in = tau_in; j.fext = fext;
   j.out = tau_out; j.qdd_out = qdd_out; j.M = M;
   j.flags = layout;
   return run_host(h, j);
}

int mecano_b200_kernel_info_get(mecano_b200_handle *h, int algo, int64_t n_states, mecano_b200_kernel_info *info)
{
   if (!h || !info || algo < 0 || algo >= MB_NUM_ALGOS) return MECANO_B200_ERR_INVALID_ARGUMENT;
   const bool warp = algo != MB_CORIOLIS && (h->variant == MECANO_B200_VARIANT_WARP || (h->variant == MECANO_B200_VARIANT_AUTO && h->warp_ok && n_states > 0 && n_states < h->warp_below[algo]));
   if (warp)
   {
      std::memset(info, 0, sizeof *info);
      cudaFuncAttributes fa;
      const bool team = h->tree.prog[algo].nb > 32;
      if (team) MB_CUDA(h, mb::team_kernel_attributes(algo, false, &fa));
      else MB_CUDA(h, mb::warp_kernel_attributes(algo, false, &fa));
      info->variant = MECANO_B200_VARIANT_WARP;
      info->block_threads = team ? mb::team_threads(h->tree.prog[algo]) : 128;
      info->states_per_block = team ? 1 : 4;
      info->regs_per_thread = fa.numRegs;
      info->local_bytes_per_thread = (int32_t)fa.localSizeBytes;
      info->static_smem_bytes = (int32_t)fa.sharedSizeBytes;
      info->sm_count = h->sm_count;
      info->max_depth = h->tree.prog[algo].max_depth;
      const double nq = h->tree.nq, nv = h->tree.nv;
      info->bytes_per_state = algo == MB_CRBA ? 8.0 * (nq + nv * nv) : 8.0 * (nq + 3.0 * nv);
      return MECANO_B200_OK;
   }
   const mb::LaunchPlan &p = h->plan[algo];
   const MbProgram &P = h->tree.prog[algo];
   std::memset(info, 0, sizeof *info);
   info->variant = MECANO_B200_VARIANT_THREAD;
   const mb::SpecKernel &sk = h->spec[algo];
   if (sk.ready())
   {
      info->block_threads = sk.opt.block;
      info->states_per_block = sk.opt.block;
      info->regs_per_thread = sk.regs;
      info->static_smem_bytes = sk.static_smem;
      info->dynamic_smem_bytes = (int32_t)sk.smem;
      info->local_bytes_per_thread = sk.local_bytes;
      info->blocks_per_sm = sk.blocks_per_sm;
      info->specialized = 1 + (sk.from_cache ? 1 : 0);
      info->tmem_stack_slots = sk.opt.tm;
      info->jit_seconds = sk.compile_seconds;
   }
   else
   {
      info->block_threads = p.block;
      info->states_per_block = p.block;
      info->regs_per_thread = p.regs;
      info->static_smem_bytes = p.static_smem;
      info->dynamic_smem_bytes = (int32_t)p.smem;
      info->local_bytes_per_thread = p.local_bytes;
      info->blocks_per_sm = p.blocks_per_sm;
      info->tmem_stack_slots = p.tm;
   }
   cudaDeviceGetAttribute(&info->sm_count, cudaDevAttrMultiProcessorCount, h->device);
   info->stack_doubles = P.stack_doubles;
   info->max_depth = P.max_depth;
   const double nq = P.nq, nv = P.nv;
   info->bytes_per_state = algo == MB_CRBA ? 8.0 * (nq + nv * nv) : (algo == MB_CORIOLIS ? 8.0 * (nq + nv + 2.0 * nv * nv) : 8.0 * (nq + 3.0 * nv));
   return MECANO_B200_OK;
}

int mecano_b200_measure_fp64_peak(int device, double *tflops)
{
   if (!tflops) return MECANO_B200_ERR_INVALID_ARGUMENT;
   cudaError_t e = cudaSetDevice(device);
   if (e == cudaSuccess) e = mb::measure_fp64_peak(tflops);
   return e == cudaSuccess ? MECANO_B200_OK : (int)e;
}

int mecano_b200_measure_fp64_sustained(int device, double seconds, double *tflops)
{
   if (!tflops || !(seconds > 0.0) || seconds > 30.0) return MECANO_B200_ERR_INVALID_ARGUMENT;
   cudaError_t e = cudaSetDevice(device);
   if (e == cudaSuccess) e = mb::measure_fp64_sustained(seconds, tflops);
   return e == cudaSuccess ? MECANO_B200_OK : (int)e;
}

int mecano_b200_measure_hbm_peak(int device, double *gbs)
{
   if (!gbs) return MECANO_B200_ERR_INVALID_ARGUMENT;
   cudaError_t e = cudaSetDevice(device);
   if (e == cudaSuccess) e = mb::measure_hbm_peak(gbs);
   return e == cudaSuccess ? MECANO_B200_OK : (int)e;
}

int mecano_b200_generate_source(const mecano_b200_tree_desc *desc, int algo, int block_threads, int tmem_slots, char *buf, int64_t capacity, int64_t *needed)
{
   if (algo < 0 || algo > 2) return fail(nullptr, MECANO_B200_ERR_INVALID_ARGUMENT, "algo must be 0 (RNEA), 1 (ABA) or 2 (CRBA)");
   mb::FlatTree tree;
   std::string err;
   int rc = mb::flatten_tree(desc, tree, err);
   if (rc != MECANO_B200_OK) return fail(nullptr, rc, err);
   mb::SpecOptions opt;
   opt.block = block_threads > 0 ? block_threads : 256;
   opt.tm = tmem_slots;
   const std::string src = mb::generate_source(algo, tree, opt);
   if (needed) *needed = (int64_t)src.size() + 1;
   if (buf && capacity > 0)
   {
      const size_t n = std::min<size_t>(src.size(), (size_t)capacity - 1);
      std::memcpy(buf, src.data(), n);
      buf[n] = 0;
   }
   return MECANO_B200_OK;
}

int mecano_b200_jit_check(const mecano_b200_tree_desc *desc, int algo, int block_threads, int tmem_slots, int64_t *cubin_bytes)
{
   if (algo < 0 || algo > 2) return fail(nullptr, MECANO_B200_ERR_INVALID_ARGUMENT, "algo must be 0 (RNEA), 1 (ABA) or 2 (CRBA)");
   mb::FlatTree tree;
   std::string err;
   int rc = mb::flatten_tree(desc, tree, err);
   if (rc != MECANO_B200_OK) return fail(nullptr, rc, err);
   mb::SpecOptions opt;
   opt.block = block_threads > 0 ? block_threads : 256;
   opt.tm = tmem_slots;
   size_t bytes = 0;
   rc = mb::spec_compile_only(mb::generate_source(algo, tree, opt), &bytes, err);
   if (cubin_bytes) *cubin_bytes = (int64_t)bytes;
   if (rc != MECANO_B200_OK) return fail(nullptr, rc, err);
   return MECANO_B200_OK;
}

int mecano_b200_host_alloc(void **ptr, int64_t bytes)
{
   if (!ptr || bytes < 0) return MECANO_B200_ERR_INVALID_ARGUMENT;
   cudaError_t e = cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault);
   return e == cudaSuccess ? MECANO_B200_OK : (int)e;
}

int mecano_b200_host_free(void *ptr)
{
   cudaError_t e = cudaFreeHost(ptr);
   return e == cudaSuccess ? MECANO_B200_OK : (int)e;
}

// ------------------------------------------------------------------------------------------------ several devices, one call
#pragma GCC visibility pop
} // extern "C"

struct mecano_b200_multi
{
   std::vector<mecano_b200_handle *> lanes;
   std::string error;
};

namespace
{
// slice of lane i: contiguous, multiples of 256 states (aligned rows on the device side), the remainder on the last lanes
void multi_slice(int64_t n, int lanes, int i, int64_t *start, int64_t *count)
{
   const int64_t blocks = (n + 255) / 256, base = blocks / lanes, extra = blocks % lanes;
   const int64_t b0 = (int64_t)i * base + std::min<int64_t>(i, extra), b1 = b0 + base + (i < extra ? 1 : 0);
   *start = std::min(n, b0 * 256);
   *count = std::min(n, b1 * 256) - *start;
}

int multi_fail(mecano_b200_multi *m, int rc, int lane)
{
   if (rc && m && lane >= 0 && lane < (int)m->lanes.size())
      m->error = "device entry " + std::to_string(lane) + " (cuda:" + std::to_string(m->lanes[(size_t)lane]->device) + "): " + m->lanes[(size_t)lane]->error;
   return rc;
}

// one job per lane from a whole-batch job: pointers offset to the lane's first state (state-major mass matrices by whole states)
int multi_run(mecano_b200_multi *m, const HostJob &whole)
{
   if (!m) return MECANO_B200_ERR_INVALID_ARGUMENT;
   if (whole.n < 0 || whole.ld < whole.n)
   {
      m->error = "n_states must be >= 0 and ld >= n_states";
      return MECANO_B200_ERR_SHAPE;
   }
   const int L = (int)m->lanes.size();
   std::vector<HostJob> jobs((size_t)L, whole);
   for (int i = 0; i < L; i++)
   {
      int64_t s0 = 0, cnt = 0;
      multi_slice(whole.n, L, i, &s0, &cnt);
      HostJob &j = jobs[(size_t)i];
      j.n = cnt;
      j.variant_n = whole.n;
      auto off = [&](const double *p) { return p ? p + s0 : nullptr; };
      auto offw = [&](double *p) { return p ? p + s0 : nullptr; };
      j.q = off(whole.q); j.qd = off(whole.qd); j.x = off(whole.x); j.fext = off(whole.fext); j.x2 = off(whole.x2); j.tau_in = off(whole.tau_in);
      j.body_acc = offw(whole.body_acc); j.joint_wrench = offw(whole.joint_wrench); j.tau_out = offw(whole.tau_out); j.qdd_out = offw(whole.qdd_out);
      const size_t nv = (size_t)m->lanes[(size_t)i]->tree.nv;
      const bool sm = (whole.flags & MECANO_B200_CRBA_STATE_MAJOR) != 0;
      if (whole.algo == MB_CRBA)
         j.out = whole.out ? (sm ? whole.out + (size_t)s0 * nv * nv : whole.out + s0) : nullptr;
      else
         j.out = offw(whole.out);
      if (whole.algo == MB_STEP)
         j.M = whole.M ? (sm ? whole.M + (size_t)s0 * nv * nv : whole.M + s0) : nullptr;
   }
   int failed = 0;
   const int rc = run_host_lanes(m->lanes, jobs, &failed);
   return multi_fail(m, rc, failed);
}
} // namespace

extern "C" {
#pragma GCC visibility push(default)

int mecano_b200_multi_create(const mecano_b200_tree_desc *desc, const int32_t *devices, int n_devices, mecano_b200_multi **out)
{
   if (!out) return fail(nullptr, MECANO_B200_ERR_INVALID_ARGUMENT, "out pointer is NULL");
   *out = nullptr;
   if (!devices || n_devices <= 0 || n_devices > 64) return fail(nullptr, MECANO_B200_ERR_INVALID_ARGUMENT, "device list must hold 1 .. 64 entries");
   mecano_b200_multi *m = new mecano_b200_multi();
   for (int i = 0; i < n_devices; i++)
   {
      mecano_b200_handle *h = nullptr;
      const int rc = mecano_b200_create(desc, devices[i], &h);
      if (rc != MECANO_B200_OK)
      {
         mecano_b200_multi_destroy(m);
         return rc; // mecano_b200_last_error(NULL) holds the message of the failed create
      }
      m->lanes.push_back(h);
   }
   *out = m;
   return MECANO_B200_OK;
}

void mecano_b200_multi_destroy(mecano_b200_multi *m)
{
   if (!m) return;
   for (mecano_b200_handle *h : m->lanes) mecano_b200_destroy(h);
   delete m;
}

const char *mecano_b200_multi_last_error(const mecano_b200_multi *m) { return m ? m->error.c_str() : mecano_b200_last_error(nullptr); }
int mecano_b200_multi_size(const mecano_b200_multi *m) { return m ? (int)m->lanes.size() : -1; }
mecano_b200_handle *mecano_b200_multi_handle(mecano_b200_multi *m, int i) { return (m && i >= 0 && i < (int)m->lanes.size()) ? m->lanes[(size_t)i] : nullptr; }

int mecano_b200_multi_slice(const mecano_b200_multi *m, int64_t n_states, int i, int64_t *start, int64_t *count)
{
   if (!m || i < 0 || i >= (int)m->lanes.size() || n_states < 0 || !start || !count) return MECANO_B200_ERR_INVALID_ARGUMENT;
   multi_slice(n_states, (int)m->lanes.size(), i, start, count);
   return MECANO_B200_OK;
}

int mecano_b200_multi_set_gravity(mecano_b200_multi *m, double gx, double gy, double gz)
{
   if (!m) return MECANO_B200_ERR_INVALID_ARGUMENT;
   for (mecano_b200_handle *h : m->lanes) mecano_b200_set_gravity(h, gx, gy, gz);
   return MECANO_B200_OK;
}

int mecano_b200_multi_rnea_host(mecano_b200_multi *m, int64_t n, int64_t ld, const double *q, const double *qd, const double *qdd, const double *fext,
                                double *tau, uint32_t flags)
{
   HostJob j;
   j.algo = MB_RNEA; j.n = n; j.ld = ld; j.q = q; j.qd = qd; j.x = qdd; j.fext = fext; j.out = tau; j.flags = flags;
   return multi_run(m, j);
}

int mecano_b200_multi_aba_host(mecano_b200_multi *m, int64_t n, int64_t ld, const double *q, const double *qd, const double *tau, const double *fext,
                               double *qdd, uint32_t flags)
{
   HostJob j;
   j.algo = MB_ABA; j.n = n; j.ld = ld; j.q = q; j.qd = qd; j.x = tau; j.fext = fext; j.out = qdd; j.flags = flags;
   return multi_run(m, j);
}

int mecano_b200_multi_crba_host(mecano_b200_multi *m, int64_t n, int64_t ld, const double *q, double *M, uint32_t layout)
{
   HostJob j;
   j.algo = MB_CRBA; j.n = n; j.ld = ld; j.q = q; j.out = M; j.flags = layout;
   return multi_run(m, j);
}

int mecano_b200_multi_step_host(mecano_b200_multi *m, int64_t n, int64_t ld, const double *q, const double *qd, const double *qdd_in, const double *tau_in,
                                const double *fext, double *tau_out, double *qdd_out, double *M, uint32_t layout)
{
   HostJob j;
   j.algo = MB_STEP; j.n = n; j.ld = ld;
   j.q = q; j.qd = qd; j.x = qdd_in; j.tau_in = tau_in; j.fext = fext;
   j.out = tau_out; j.qdd_out = qdd_out; j.M = M;
   j.flags = layout;
   return multi_run(m, j);
}

#pragma GCC visibility pop
} // extern "C"
