// Runtime of the SIMT emulator (see cuda_emu.h).  Test infrastructure only.
#include "cuda_emu.h"

#include <mutex>

namespace emu {
thread_local Block* tb = nullptr;
thread_local uint3_ tl_threadIdx;

static void fiber_entry() {
  Block* b = tb;
  (*b->body)();
  b->done[b->cur] = true;
  b->exited++;
  b->warp_exited[b->cur >> 5]++;
  // a thread that leaves releases barriers the remaining threads are waiting on
  if (b->bar_count > 0 && b->bar_count >= b->nthreads - b->exited) { b->bar_count = 0; b->bar_gen++; }
  Warp& w = b->warps[b->cur >> 5];
  int width = std::min(32, b->nthreads - (b->cur >> 5) * 32) - b->warp_exited[b->cur >> 5];
  if (w.count > 0 && w.count >= width) { w.count = 0; w.gen++; }
  swapcontext(&b->ctx[b->cur], &b->main_ctx);
}

static void run_block(Block* b, dim3 bidx) {
  b->bidx = bidx;
  b->exited = 0; b->bar_gen = 0; b->bar_count = 0;
  int n = b->nthreads;
  for (int w = 0; w < (n + 31) / 32; ++w) { b->warps[w].gen = 0; b->warps[w].count = 0; b->warp_exited[w] = 0; }
  for (int t = 0; t < n; ++t) {
    b->done[t] = false;
    getcontext(&b->ctx[t]);
    b->ctx[t].uc_stack.ss_sp = b->stacks + (size_t)t * kStack;
    b->ctx[t].uc_stack.ss_size = kStack;
    b->ctx[t].uc_link = &b->main_ctx;
    makecontext(&b->ctx[t], fiber_entry, 0);
  }
  while (b->exited < n) {
    for (int t = 0; t < n; ++t) {
      if (b->done[t]) continue;
      b->cur = t;
      unsigned bx = b->bdim.x, by = b->bdim.y;
      tl_threadIdx.x = t % bx; tl_threadIdx.y = (t / bx) % by; tl_threadIdx.z = t / (bx * by);
      swapcontext(&b->main_ctx, &b->ctx[t]);
    }
  }
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  int nthreads = block.x * block.y * block.z;
  if (nthreads <= 0 || nthreads > kMaxThreads) { fprintf(stderr, "emu: bad block size %d\n", nthreads); abort(); }
  size_t nblocks = (size_t)grid.x * grid.y * grid.z;
  if (nblocks == 0) return;
  static int nworkers = [] { const char* e = getenv("HULC_EMU_THREADS"); int n = e ? atoi(e) : (int)std::thread::hardware_concurrency(); return std::max(1, n); }();
  int nw = (int)std::min<size_t>(nworkers, nblocks);
  std::atomic<size_t> next{0};
  static std::mutex pool_mu;
  static std::vector<Block*> pool;
  auto worker = [&]() {
    Block* b = nullptr;
    {
      std::lock_guard<std::mutex> g(pool_mu);
      if (!pool.empty()) { b = pool.back(); pool.pop_back(); }
    }
    if (!b) { b = new Block(); b->stacks = (char*)aligned_alloc(4096, kStack * kMaxThreads); }
    tb = b;
    b->nthreads = nthreads; b->bdim = block; b->gdim = grid; b->body = &body;
    b->dyn_smem.assign(smem + 16, 0);
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= nblocks) break;
      dim3 bi((unsigned)(i % grid.x), (unsigned)((i / grid.x) % grid.y), (unsigned)(i / ((size_t)grid.x * grid.y)));
      run_block(b, bi);
    }
    tb = nullptr;
    std::lock_guard<std::mutex> g(pool_mu);
    pool.push_back(b);
  };
  if (nw == 1) { worker(); return; }
  std::vector<std::thread> ts;
  for (int i = 0; i < nw; ++i) ts.emplace_back(worker);
  for (auto& t : ts) t.join();
}
}  // namespace emu
