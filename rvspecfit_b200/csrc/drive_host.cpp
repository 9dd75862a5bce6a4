// Native round loop of a lock-step Nelder-Mead stage of the batched fit.
//
// One round of the stage = ask the stepper for the next trial points (rvs_nm_request),
// turn them into the evaluation call's pinned upload buffers (rvs_fit_pack), launch the
// call's captured CUDA graph on the slot's stream, wait for its event, reduce the
// downloads to objective values (rvs_fit_collect) and feed them back (rvs_nm_feed).
// Driven from Python this costs 0.3-0.5 ms of interpreter time per round -- more than
// the device needs for a call of a few hundred items, so the latency-bound tails of the
// optimiser (few live problems) and every concurrent lock-step set queued behind the
// interpreter lock.  rvs_nm_drive runs the rounds here, without the interpreter, and
// returns to the caller only when it needs something only the caller can do:
//   RVS_DRIVE_DONE       every problem has stopped
//   RVS_DRIVE_PEEL       `stop_stopped` problems have stopped (the caller hands them on)
//   RVS_DRIVE_LAUNCH     no captured graph for this call's configuration: the caller
//                        launches it its own way (and captures it on the second sighting)
//   RVS_DRIVE_REDO       the fused path could not settle some items (f_redo): the caller
//                        evaluates them through the general path and patches f_out
//   RVS_DRIVE_PYEVAL     the call does not fit the fused path at all (vsini bound): the
//                        caller evaluates the whole request and writes f_out
// and is called again to go on.  The decision rules, the packing and the reduction are
// the very functions the Python route uses, so both routes visit the same points.
#include <cuda_runtime_api.h>
#include <math.h>
#include <sched.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/rvs_b200.h"

namespace {

struct Drive {
  std::vector<int32_t> idx, obj;
  std::vector<double> X, logv;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<int> logcols;
};

}  // namespace

extern "C" int64_t rvs_fit_round_items(int64_t K) {
  // Launch configurations of an evaluation call: four per octave (16-item granules), so
  // that an optimiser whose live set shrinks call by call keeps hitting a few dozen
  // captured graphs.  Items K..Kp-1 are absent on every arm.
  if (K <= 16) return 16;
  int64_t p = 16;
  while (p * 2 < K) p *= 2;          // largest power of two below K
  int64_t step = p / 4 < 16 ? 16 : p / 4;
  return (K + step - 1) / step * step;
}

extern "C" void *rvs_stream_create(int high_priority) {
  // A stream of its own for every in-flight evaluation: framework stream pools hand the
  // same few handles out again and again, and two evaluations that share a stream cannot
  // be captured and launched by different threads.
  // high_priority: 0 least (that of the default stream), 1 halfway, 2 greatest
  int least = 0, greatest = 0;
  cudaStream_t st = nullptr;
  if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) return nullptr;
  const int prio = high_priority <= 0 ? least
                   : (high_priority == 1 ? (least + greatest) / 2 : greatest);
  if (cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio) != cudaSuccess) return nullptr;
  return st;
}

extern "C" void rvs_stream_destroy(void *st) {
  if (st) cudaStreamDestroy(static_cast<cudaStream_t>(st));
}

extern "C" void *rvs_drive_create(int64_t cap, int N) {
  if (cap < 1 || N < 1) return nullptr;
  Drive *d = new Drive;
  d->idx.resize(cap);
  d->obj.resize(cap);
  d->X.resize((size_t)cap * N);
  d->logv.resize((size_t)cap * N);
  if (cudaEventCreate(&d->ev0) != cudaSuccess || cudaEventCreate(&d->ev1) != cudaSuccess) {
    delete d;
    return nullptr;
  }
  return d;
}

extern "C" void rvs_drive_destroy(void *h) {
  Drive *d = static_cast<Drive *>(h);
  if (!d) return;
  if (d->ev0) cudaEventDestroy(d->ev0);
  if (d->ev1) cudaEventDestroy(d->ev1);
  delete d;
}

extern "C" int rvs_drive_request(void *h, const int32_t **idx, const double **X, const int32_t **obj) {
  Drive *d = static_cast<Drive *>(h);
  if (!d) return RVS_E_ARG;
  if (idx) *idx = d->idx.data();
  if (X) *X = d->X.data();
  if (obj) *obj = d->obj.data();
  return 0;
}

namespace {
// the optimiser behind the loop: Nelder-Mead (nm_host.cpp) or BFGS (bfgs_host.cpp)
struct StepOps {
  int64_t (*request)(void *, int, int32_t *, double *, int64_t);
  int (*feed)(void *, const double *, int64_t);
  int64_t (*live)(void *, uint8_t *);
};
int64_t bfgs_request(void *h, int, int32_t *idx, double *X, int64_t cap) {
  return rvs_bfgs_request(h, idx, X, cap);
}
int drive_loop(const StepOps &ops, void *nm, void *drive, const rvs_fit_layout *lay, rvs_drive *io);
}  // namespace

extern "C" int rvs_nm_drive(void *nm, void *drive, const rvs_fit_layout *lay, rvs_drive *io) {
  static const StepOps ops = {rvs_nm_request, rvs_nm_feed, rvs_nm_live};
  return drive_loop(ops, nm, drive, lay, io);
}

extern "C" int rvs_bfgs_drive(void *bfgs, void *drive, const rvs_fit_layout *lay, rvs_drive *io) {
  static const StepOps ops = {bfgs_request, rvs_bfgs_feed, rvs_bfgs_live};
  return drive_loop(ops, bfgs, drive, lay, io);
}

namespace {
int drive_loop(const StepOps &ops, void *nm, void *drive, const rvs_fit_layout *lay, rvs_drive *io) {
  Drive *d = static_cast<Drive *>(drive);
  if (!nm || !d || !lay || !io || !io->objmap) return RVS_E_ARG;
  const int N = lay->nfit, ns = lay->nspec, narm = lay->narm;
  const int64_t cap = (int64_t)d->idx.size();
  // columns of the fitted vector that hold log-mapped parameters, in parameter order
  d->logcols.clear();
  {
    int pos = 1 + (lay->fit_vsini ? 1 : 0);
    for (int j = 0; j < ns; j++)
      if (!(lay->fixmask >> j & 1)) {
        if (lay->logmask >> j & 1) d->logcols.push_back(pos);
        pos++;
      }
  }
  cudaStream_t stream = static_cast<cudaStream_t>(io->stream);
  for (;;) {
    if (io->state == RVS_DRIVE_IDLE) {
      const int64_t n = ops.request(nm, io->speculate_below, d->idx.data(), d->X.data(), cap);
      if (n == 0) return RVS_DRIVE_DONE;
      if (n > cap || n > io->cap) return RVS_E_LIMIT;
      const int64_t Kp = rvs_fit_round_items(n);
      if (Kp > io->cap) return RVS_E_LIMIT;
      for (int64_t k = 0; k < n; k++) d->obj[k] = io->objmap[d->idx[k]];
      const size_t nl = d->logcols.size();
      for (size_t r = 0; r < nl; r++) {
        const int c = d->logcols[r];
        double *row = d->logv.data() + r * (size_t)n;
        for (int64_t k = 0; k < n; k++) row[k] = log10(d->X[(size_t)k * N + c]);
      }
      double vmax = 0.0;
      const int rc = rvs_fit_pack(lay, n, Kp, d->obj.data(), d->X.data(), nl ? d->logv.data() : nullptr,
                                  io->h_in, io->h_oix, io->f_prior, io->f_pen, io->f_wall, &vmax);
      if (rc) return rc;
      if (vmax > 0) {     // tap bound of the call in coarse steps (LikelihoodEngine._tap_bound)
        double rounded = exp2(ceil(log2(vmax)));
        if (rounded < 16.0) rounded = 16.0;
        if (rounded <= io->fused_vmax) vmax = rounded;
      }
      io->K = n;
      io->Kp = Kp;
      io->vmax = vmax;
      io->state = RVS_DRIVE_PACKED;
      if (vmax > io->fused_vmax) return RVS_DRIVE_PYEVAL;
    }
    if (io->state == RVS_DRIVE_PACKED) {
      void *exec = nullptr;
      int nk = 0;
      for (int g = 0; g < io->ngraph; g++)
        if (io->g_kp[g] == io->Kp && io->g_vmax[g] == io->vmax) {
          exec = io->g_exec[g];
          nk = io->g_nk[g];
          break;
        }
      if (!exec) return RVS_DRIVE_LAUNCH;
      if (io->epoch_event) cudaEventRecord(d->ev0, stream);
      if (cudaGraphLaunch(static_cast<cudaGraphExec_t>(exec), stream) != cudaSuccess) return RVS_E_CUDA;
      io->graph_launches += 1;
      io->graph_kernels += nk;
      io->timed = io->epoch_event != nullptr;
      io->state = RVS_DRIVE_LAUNCHED;
    }
    if (io->state == RVS_DRIVE_LAUNCHED) {
      // (after RVS_DRIVE_LAUNCH the caller's launch is on the stream: the event follows it)
      if (cudaEventRecord(d->ev1, stream) != cudaSuccess) return RVS_E_CUDA;
      // poll and yield: as quick as a spinning wait when the thread has a core of its
      // own, and out of the way of the other sets' host threads when it has not (a rank of
      // an 8-GPU job has four cores for three such loops and the interpreter)
      for (;;) {
        const cudaError_t e = cudaEventQuery(d->ev1);
        if (e == cudaSuccess) break;
        (void)cudaGetLastError();
        if (e != cudaErrorNotReady) return RVS_E_CUDA;
        sched_yield();
      }
      if (io->timed && io->t_rec && io->t_n < io->t_cap) {
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, static_cast<cudaEvent_t>(io->epoch_event), d->ev0);
        cudaEventElapsedTime(&b, static_cast<cudaEvent_t>(io->epoch_event), d->ev1);
        double *r = io->t_rec + 3 * io->t_n++;
        r[0] = a; r[1] = b; r[2] = (double)io->K;
      }
      io->timed = 0;
      io->h2d_bytes += (2 + ns) * io->Kp * 8 + (int64_t)narm * io->Kp * 4;
      io->d2h_bytes += 2 * (int64_t)narm * io->Kp * 12;
      const int64_t nredo = rvs_fit_collect(lay, io->K, io->Kp, d->obj.data(), io->h_in, io->h_chi,
                                            io->h_flags, io->shared_locate, 1, io->f_prior, io->f_pen,
                                            io->f_wall, io->f_out, io->f_redo);
      if (nredo < 0) return (int)nredo;
      io->state = RVS_DRIVE_COLLECTED;
      if (nredo > 0) return RVS_DRIVE_REDO;
    }
    if (io->state == RVS_DRIVE_COLLECTED) {
      const int rc = ops.feed(nm, io->f_out, io->K);
      if (rc) return rc;
      io->rounds += 1;
      io->items += io->K;
      io->state = RVS_DRIVE_IDLE;
      if (io->stop_stopped > 0) {
        const int64_t live = ops.live(nm, nullptr);
        if (live > 0 && io->nprob - live >= io->stop_stopped) return RVS_DRIVE_PEEL;
      }
    }
  }
}
}  // namespace
