// TEST INFRASTRUCTURE ONLY - a fiber-based SIMT emulator.
//
// Runs the warp-level solver source (mpc_racing_b200/csrc/mpcr_ipm.cuh) on the
// CPU so that the kernel LOGIC can be compared with the oracle in a container
// without a GPU.  32 ucontext fibers play the lanes of one warp; every warp
// collective (shuffle, syncwarp) deposits its operand and yields round-robin,
// so a lane resumes only after all other lanes reached the same collective.
// The product never links this file: libmpcr.so is built from the .cu sources
// only and has no CPU path.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

namespace emu {

constexpr int kLanes = 32;
constexpr size_t kStackBytes = 1 << 20;

struct Warp {
  ucontext_t main_ctx;
  ucontext_t ctx[kLanes];
  char* stacks;
  bool done[kLanes];
  int cur;
  uint64_t xbuf[2][kLanes];
  unsigned cnt[kLanes];
  void (*body)(void*);
  void* arg;
};

inline Warp*& current() {
  static thread_local Warp* w = nullptr;
  return w;
}

inline int lane() { return current()->cur; }

inline void yield_next() {
  Warp* w = current();
  int me = w->cur;
  int nxt = me;
  for (int i = 0; i < kLanes; ++i) {
    nxt = (nxt + 1) % kLanes;
    if (!w->done[nxt]) break;
  }
  if (nxt == me) return;
  w->cur = nxt;
  swapcontext(&w->ctx[me], &w->ctx[nxt]);
}

inline uint64_t shfl_u64(uint64_t v, int src) {
  Warp* w = current();
  int me = w->cur;
  unsigned par = w->cnt[me] & 1u;
  w->xbuf[par][me] = v;
  w->cnt[me]++;
  yield_next();
  return w->xbuf[par][src & (kLanes - 1)];
}

inline double shfl_f64(double v, int src) {
  uint64_t b;
  memcpy(&b, &v, 8);
  b = shfl_u64(b, src);
  double r;
  memcpy(&r, &b, 8);
  return r;
}
inline int shfl_i32(int v, int src) { return (int)(int64_t)shfl_u64((uint64_t)(int64_t)v, src); }
inline void syncwarp() { (void)shfl_u64(0, 0); }

inline void fiber_entry() {
  Warp* w = current();
  int me = w->cur;
  w->body(w->arg);
  w->done[me] = true;
  // hand over to the next unfinished lane, or back to the launcher
  int nxt = -1;
  for (int i = 1; i <= kLanes; ++i) {
    int c = (me + i) % kLanes;
    if (!w->done[c]) { nxt = c; break; }
  }
  if (nxt < 0) {
    setcontext(&w->main_ctx);
  } else {
    w->cur = nxt;
    setcontext(&w->ctx[nxt]);
  }
}

// Run body(arg) once per lane, in lock-step at collectives.
inline void run_warp(void (*body)(void*), void* arg) {
  Warp* w = (Warp*)calloc(1, sizeof(Warp));
  w->stacks = (char*)malloc(kStackBytes * kLanes);
  w->body = body;
  w->arg = arg;
  Warp* saved = current();
  current() = w;
  for (int i = 0; i < kLanes; ++i) {
    getcontext(&w->ctx[i]);
    w->ctx[i].uc_stack.ss_sp = w->stacks + (size_t)i * kStackBytes;
    w->ctx[i].uc_stack.ss_size = kStackBytes;
    w->ctx[i].uc_link = nullptr;
    makecontext(&w->ctx[i], (void (*)())fiber_entry, 0);
  }
  w->cur = 0;
  swapcontext(&w->main_ctx, &w->ctx[0]);
  current() = saved;
  free(w->stacks);
  free(w);
}

}  // namespace emu
