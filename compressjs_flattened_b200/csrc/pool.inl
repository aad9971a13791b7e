// pool.inl -- the shard scheduler (included by bz2b200.cu inside its namespace).
//
// One bzip2 stream is a sequence of SHARDS (slices of the input, each followed by a halo).  A shard is compressed by
// one LANE (a context with its own stream on some device); the lanes of a pool -- two per device by default -- work on
// consecutive shards at the same time, so inside ONE blocking call
//   * the host->device copy of shard k+1 runs under the kernels of shard k (a staging thread copies pageable input
//     through page-locked memory first),
//   * the device->host copy of segment k runs under the kernels of the other lane,
//   * several devices (in this process, or one rank per device in other processes) take every n-th shard.
// What travels from shard j-1 to shard j are scalars (no collective on the data path):
//   cut  : the input offset at which shard j's first block starts (known after the cut walk of shard j-1),
//   tail : end bit of segment j-1, blocks so far, CRC fold so far (known after shard j-1's stages).
// They go through an Exchange: process memory (mutex + condition variable) for the lanes of one process, a page of
// POSIX shared memory for ranks of one node (bz2b200_group_*).  Every wait has a deadline and a failed shard wakes
// everybody, so a dead peer is an error code, not a hang.

struct ExTail { u64 end_bit = 32, blocks = 0; u32 fold = 0; u64 aux = 0; };

struct Exchange {
  virtual ~Exchange() {}
  virtual int put_cut(int j, u64 next_start, u64 aux = 0) = 0;
  virtual int get_cut(int j, u64 *next_start, u64 *aux = nullptr) = 0;
  virtual int put_tail(int j, const ExTail &t) = 0;
  virtual int get_tail(int j, ExTail *t) = 0;
  virtual void fail(int rc) = 0;
};

// ---- lanes of one process ----
struct LocalExchange : Exchange {
  std::mutex mu;
  std::condition_variable cv;
  std::vector<u64> cut, cut_aux;
  std::vector<ExTail> tail;
  std::vector<u8> cut_ok, tail_ok;
  int err = 0;
  int timeout_ms = 600000;
  explicit LocalExchange(int n) : cut((size_t)n), cut_aux((size_t)n), tail((size_t)n), cut_ok((size_t)n, 0), tail_ok((size_t)n, 0) {}
  int put_cut(int j, u64 v, u64 aux = 0) override {
    { std::lock_guard<std::mutex> g(mu); cut[(size_t)j] = v; cut_aux[(size_t)j] = aux; cut_ok[(size_t)j] = 1; }
    cv.notify_all();
    return 0;
  }
  int get_cut(int j, u64 *v, u64 *aux = nullptr) override {
    std::unique_lock<std::mutex> l(mu);
    if (!cv.wait_for(l, std::chrono::milliseconds(timeout_ms), [&] { return err || cut_ok[(size_t)j]; })) return BZ2B200_E_PEER;
    if (err) return err;
    *v = cut[(size_t)j];
    if (aux) *aux = cut_aux[(size_t)j];
    return 0;
  }
  int put_tail(int j, const ExTail &t) override {
    { std::lock_guard<std::mutex> g(mu); tail[(size_t)j] = t; tail_ok[(size_t)j] = 1; }
    cv.notify_all();
    return 0;
  }
  int get_tail(int j, ExTail *t) override {
    std::unique_lock<std::mutex> l(mu);
    if (!cv.wait_for(l, std::chrono::milliseconds(timeout_ms), [&] { return err || tail_ok[(size_t)j]; })) return BZ2B200_E_PEER;
    if (err) return err;
    *t = tail[(size_t)j];
    return 0;
  }
  void fail(int rc) override {
    { std::lock_guard<std::mutex> g(mu); if (!err) err = rc ? rc : BZ2B200_E_PEER; }
    cv.notify_all();
  }
};

// ---- ranks of one node: a file in /dev/shm ----
// Slot j holds what shard j publishes, stamped with the EPOCH (the number of the collective call, counted alike by every
// rank), so nothing is ever cleared.  A call ends with a barrier (arrive counter), hence no rank is more than one call
// ahead and a slot is never overwritten under a reader.
#define GRP_MAGIC 0x62327a4752503031ull  // "b2zGRP01"
#define GRP_SLOTS 4096
struct GrpSlot {
  std::atomic<u64> cut_seq, tail_seq;
  u64 next_start, cut_aux, end_bit, blocks, aux;
  u32 fold, pad;
};
struct GrpHeader {
  std::atomic<u64> magic;
  std::atomic<u64> attached, arrive, err_epoch;
  std::atomic<int> err_code;
  u32 world;
  GrpSlot slot[GRP_SLOTS];
};
struct Group {
  GrpHeader *h = nullptr;
  int rank = 0, world = 1, fd = -1;
  u64 epoch = 0;
  int timeout_ms = 120000;
  std::string path;
  bool creator = false;
};
struct GroupExchange : Exchange {
  Group *g;
  explicit GroupExchange(Group *g_) : g(g_) {}
  template <typename F> int wait(F ready) {
    auto t0 = std::chrono::steady_clock::now();
    for (unsigned spin = 0;; spin++) {
      if (ready()) return 0;
      if (g->h->err_epoch.load(std::memory_order_acquire) == g->epoch) { int e = g->h->err_code.load(); return e ? e : BZ2B200_E_PEER; }
      if ((spin & 63) == 63) {
        if (std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count() > g->timeout_ms) {
          fail(BZ2B200_E_PEER);
          return BZ2B200_E_PEER;
        }
        std::this_thread::yield();
      }
    }
  }
  int put_cut(int j, u64 v, u64 aux = 0) override {
    if (j < 0 || j >= GRP_SLOTS) return BZ2B200_E_ARG;
    g->h->slot[j].next_start = v;
    g->h->slot[j].cut_aux = aux;
    g->h->slot[j].cut_seq.store(g->epoch, std::memory_order_release);
    return 0;
  }
  int get_cut(int j, u64 *v, u64 *aux = nullptr) override {
    if (j < 0 || j >= GRP_SLOTS) return BZ2B200_E_ARG;
    int rc = wait([&] { return g->h->slot[j].cut_seq.load(std::memory_order_acquire) == g->epoch; });
    if (!rc) { *v = g->h->slot[j].next_start; if (aux) *aux = g->h->slot[j].cut_aux; }
    return rc;
  }
  int put_tail(int j, const ExTail &t) override {
    if (j < 0 || j >= GRP_SLOTS) return BZ2B200_E_ARG;
    GrpSlot &s = g->h->slot[j];
    s.end_bit = t.end_bit; s.blocks = t.blocks; s.fold = t.fold; s.aux = t.aux;
    s.tail_seq.store(g->epoch, std::memory_order_release);
    return 0;
  }
  int get_tail(int j, ExTail *t) override {
    if (j < 0 || j >= GRP_SLOTS) return BZ2B200_E_ARG;
    GrpSlot &s = g->h->slot[j];
    int rc = wait([&] { return s.tail_seq.load(std::memory_order_acquire) == g->epoch; });
    if (!rc) { t->end_bit = s.end_bit; t->blocks = s.blocks; t->fold = s.fold; t->aux = s.aux; }
    return rc;
  }
  void fail(int rc) override {
    g->h->err_code.store(rc ? rc : BZ2B200_E_PEER);
    g->h->err_epoch.store(g->epoch, std::memory_order_release);
  }
  // end of a collective call: nobody starts the next epoch before everybody has left this one
  int barrier() {
    g->h->arrive.fetch_add(1, std::memory_order_acq_rel);
    const u64 want = g->epoch * (u64)g->world;
    return wait([&] { return g->h->arrive.load(std::memory_order_acquire) >= want; });
  }
};

int group_open(const char *name, int rank, int world, int timeout_ms, Group **out) {
  if (!name || !*name || rank < 0 || world < 1 || rank >= world || world > GRP_SLOTS) return BZ2B200_E_ARG;
  Group *g = new Group();
  g->rank = rank; g->world = world;
  if (timeout_ms > 0) g->timeout_ms = timeout_ms;
  g->path = std::string("/dev/shm/bz2b200_grp_") + name;
  for (size_t i = strlen("/dev/shm/"); i < g->path.size(); i++)  // the name stays a file name inside /dev/shm
    if (g->path[i] == ' ' || g->path[i] == '/' || g->path[i] == '.') g->path[i] = '_';
  const size_t bytes = sizeof(GrpHeader);
  // whoever creates the file (O_EXCL) sizes it; a new file is all zeros, which is a valid empty state; `magic` is set last
  int fd = open(g->path.c_str(), O_RDWR | O_CREAT | O_EXCL, 0600);
  if (fd >= 0) {
    g->creator = true;
    if (ftruncate(fd, (off_t)bytes) != 0) { close(fd); unlink(g->path.c_str()); delete g; return BZ2B200_E_PEER; }
  } else {
    auto t0 = std::chrono::steady_clock::now();
    for (;;) {
      fd = open(g->path.c_str(), O_RDWR);
      struct stat sb;
      if (fd >= 0 && fstat(fd, &sb) == 0 && (size_t)sb.st_size >= bytes) break;
      if (fd >= 0) close(fd);
      fd = -1;
      if (std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count() > g->timeout_ms) { delete g; return BZ2B200_E_PEER; }
      std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
  }
  void *m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  if (m == MAP_FAILED) { close(fd); delete g; return BZ2B200_E_PEER; }
  g->fd = fd;
  g->h = reinterpret_cast<GrpHeader *>(m);
  if (g->creator) { g->h->world = (u32)world; g->h->magic.store(GRP_MAGIC, std::memory_order_release); }
  // handshake: everybody attached (also catches a stale file of another world size)
  auto t0 = std::chrono::steady_clock::now();
  while (g->h->magic.load(std::memory_order_acquire) != GRP_MAGIC) {
    if (std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count() > g->timeout_ms) { munmap(m, bytes); close(fd); delete g; return BZ2B200_E_PEER; }
    std::this_thread::yield();
  }
  if (g->h->world != (u32)world) { munmap(m, bytes); close(fd); delete g; return BZ2B200_E_ARG; }
  g->h->attached.fetch_add(1);
  while (g->h->attached.load() < (u64)world) {
    if (std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count() > g->timeout_ms) { munmap(m, bytes); close(fd); delete g; return BZ2B200_E_PEER; }
    std::this_thread::yield();
  }
  if (g->creator) unlink(g->path.c_str());  // everybody holds a mapping now: the name can go (nothing is left behind on a crash)
  *out = g;
  return 0;
}
void group_close(Group *g) {
  if (!g) return;
  if (g->h) munmap(g->h, sizeof(GrpHeader));
  if (g->fd >= 0) close(g->fd);
  delete g;
}

// ---- lanes ----
struct Lane {
  Ctx *c = nullptr;
  cudaStream_t copy = nullptr;
  cudaEvent_t in_ev[2] = {nullptr, nullptr};
  DevBuf in_slot[2];
  DevBuf out_slot[2];                   // decode: the bytes of a range until their place in the output is known
  void *stage[2] = {nullptr, nullptr};  // page-locked staging of pageable input
  size_t stage_cap[2] = {0, 0};
  std::string err;
};
struct Pool {
  std::vector<Lane *> lanes;  // ordered lane-major: (dev0,l0) (dev1,l0) ... (dev0,l1) (dev1,l1) ...: consecutive shards meet different devices
  int n_dev = 0;
  std::string err;
  bz2b200_stats st{};
  bool force_staging = false;   // tests only: treat the input as pageable
  u32 cap_override = 0, batch_override = 0, dec_batch = 0;
  size_t halo0 = 0;             // tests only: first halo tried (0 = 5/4 of a block + 64 KiB)
  size_t plan_first = 0;        // shard plan: first shard (0 = six blocks) and growth per wave (0 = 2.5)
  double plan_growth = 0;
};

int ctx_new(int device, Ctx **out);   // bz2b200.cu
void ctx_delete(Ctx *c);

void pool_delete(Pool *p) {
  if (!p) return;
  for (Lane *L : p->lanes) {
    if (!L) continue;
    if (L->c) {
      cudaSetDevice(L->c->device);
      for (int s = 0; s < 2; s++) {
        if (L->in_slot[s].p) cudaFree(L->in_slot[s].p);
        if (L->out_slot[s].p) cudaFree(L->out_slot[s].p);
        if (L->stage[s]) cudaFreeHost(L->stage[s]);
        if (L->in_ev[s]) cudaEventDestroy(L->in_ev[s]);
      }
      if (L->copy) cudaStreamDestroy(L->copy);
      ctx_delete(L->c);
    }
    delete L;
  }
  delete p;
}
int pool_new(const int *devices, int n_dev, int lanes_per_dev, Pool **out) {
  if (!devices || n_dev < 1 || lanes_per_dev < 1 || lanes_per_dev > 4) return BZ2B200_E_ARG;
  Pool *p = new Pool();
  p->n_dev = n_dev;
  for (int l = 0; l < lanes_per_dev; l++)
    for (int d = 0; d < n_dev; d++) {
      Lane *L = new Lane();
      p->lanes.push_back(L);
      int rc = ctx_new(devices[d], &L->c);
      if (rc) { pool_delete(p); return rc; }
      if (cudaStreamCreateWithFlags(&L->copy, cudaStreamNonBlocking) != cudaSuccess) { pool_delete(p); return BZ2B200_E_CUDA; }
      for (int s = 0; s < 2; s++) if (cudaEventCreate(&L->in_ev[s]) != cudaSuccess) { pool_delete(p); return BZ2B200_E_CUDA; }
    }
  *out = p;
  return 0;
}

struct ShardJob {
  const u8 *src = nullptr;   // own_len bytes of the stream, then the halo
  size_t n_avail = 0;        // bytes to use at src (own + current halo)
  size_t n_max = 0;          // bytes readable at src (the halo may grow up to here)
  size_t own_len = 0;
  u64 base = 0;              // offset of src[0] in the whole input
  int index = 0;             // position of the shard in the stream
  bool on_device = false;
  bool last = false;         // the stream ends with this shard
};
struct ShardOut {
  u8 *seg = nullptr;         // ranks mode: page-locked segment (bz2b200_free); whole mode: nullptr (bytes go to their place)
  size_t seg_bytes = 0;
  bz2b200_shard_info info{};
  u64 bit_off = 0;
  u8 first_byte = 0;
  ExTail tail;               // running totals after this shard
};

static bool is_pageable(const void *p) {
#ifndef BZ_SIM
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return a.type == cudaMemoryTypeUnregistered;
#else
  (void)p;
  return false;
#endif
}

// ---- host placement: threads that feed a device run on the CPUs of the NUMA node the device hangs off -------------
// Page-locked memory is placed where the thread that first touches it runs, and a copy whose host side sits on the
// other socket crosses the socket interconnect: with eight ranks on a two-socket box that is the difference between
// every rank having its own PCIe root and all of them sharing one link.  The node comes from sysfs
// (/sys/bus/pci/devices/<bus id>/numa_node); a virtual machine that hides it (-1) leaves the threads where they are.
// Library-owned threads (lanes 1.., upload helpers) bind themselves; the CALLER's thread is only bound when the host
// asks for it (bz2b200_bind_thread_to_device), since narrowing a thread the library does not own is the host's call.
// BZ2B200_NUMA=0 switches all of it off.
struct NodeCpus { int node = -2; cpu_set_t set; };  // node -2: not looked up yet
static int numa_node_cpus(int device, cpu_set_t *out) {
  static std::mutex mu;
  static NodeCpus tab[64];
  if (device < 0 || device >= 64) return -1;
  std::lock_guard<std::mutex> g(mu);
  NodeCpus &e = tab[device];
  if (e.node == -2) {
    e.node = -1;
    CPU_ZERO(&e.set);
    const char *sw = getenv("BZ2B200_NUMA");
    const char *root = getenv("BZ2B200_SYSFS");  // tests point this at a made-up tree
    if (!root || !root[0]) root = "/sys";
    char bus[32] = {0};
    bool have_bus = false;
#ifndef BZ_SIM
    have_bus = cudaDeviceGetPCIBusId(bus, sizeof bus, device) == cudaSuccess;
    if (!have_bus) cudaGetLastError();
#else
    if (const char *b = getenv("BZ2B200_SIM_BUSID")) { snprintf(bus, sizeof bus, "%s", b); have_bus = true; }  // the simulator has no PCI address
#endif
    if (!(sw && sw[0] == '0') && have_bus) {
      for (char *q = bus; *q; q++) *q = (char)tolower(*q);
      char path[512];
      snprintf(path, sizeof path, "%s/bus/pci/devices/%s/numa_node", root, bus);
      int node = -1;
      if (FILE *f = fopen(path, "r")) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
      if (node >= 0) {
        snprintf(path, sizeof path, "%s/devices/system/node/node%d/cpulist", root, node);
        if (FILE *f = fopen(path, "r")) {
          int a, b, got = 0;
          while (fscanf(f, "%d", &a) == 1) {  // "0-15,32-47"
            b = a;
            int ch = fgetc(f);
            if (ch == '-') { if (fscanf(f, "%d", &b) != 1) break; ch = fgetc(f); }
            for (int k = a; k <= b && k < CPU_SETSIZE; k++) { CPU_SET(k, &e.set); got++; }
            if (ch != ',') break;
          }
          fclose(f);
          if (got) e.node = node;
        }
      }
    }
  }
  if (e.node >= 0 && out) *out = e.set;
  return e.node;
}
// binds the calling thread; returns the node, or -1 when nothing is known or the node has none of the thread's CPUs
static int numa_bind_self(int device) {
  cpu_set_t want, have, both;
  const int node = numa_node_cpus(device, &want);
  if (node < 0) return -1;
  if (sched_getaffinity(0, sizeof have, &have) != 0) return -1;
  CPU_AND(&both, &want, &have);
  if (CPU_COUNT(&both) == 0) return -1;
  if (!CPU_EQUAL(&both, &have) && sched_setaffinity(0, sizeof both, &both) != 0) return -1;
  return node;
}

// host -> device copy of one shard into an input slot of the lane (asynchronous on the lane's copy stream; pageable
// input is first copied into page-locked staging memory by the calling thread -- a helper thread, see lane_run)
static int lane_upload(Lane *L, const ShardJob &job, int slot, bool pageable, const u8 **d_in) {
  Ctx *c = L->c;
  CK(cudaSetDevice(c->device));
  if (job.on_device) { *d_in = job.src; return 0; }
  ENS(L->in_slot[slot], job.n_avail + 64);
  const void *src = job.src;
  if (pageable && job.n_avail) {
    if (L->stage_cap[slot] < job.n_avail) {
      if (L->stage[slot]) CK(cudaFreeHost(L->stage[slot]));
      L->stage[slot] = nullptr; L->stage_cap[slot] = 0;
      size_t want = job.n_avail + job.n_avail / 8 + 4096;
      CK(cudaMallocHost(&L->stage[slot], want));
      L->stage_cap[slot] = want;
    }
    memcpy(L->stage[slot], job.src, job.n_avail);
    src = L->stage[slot];
  }
  if (job.n_avail) CK(cudaMemcpyAsync(L->in_slot[slot].p, src, job.n_avail, cudaMemcpyHostToDevice, L->copy));
  CK(cudaEventRecord(L->in_ev[slot], L->copy));
  *d_in = P<u8>(L->in_slot[slot]);
  return 0;
}

struct PoolRun {
  Pool *pool;
  Exchange *ex;
  int level;
  bool pageable;
  bool keep_on_device;     // ranks mode: segments stay in HBM (device-resident timing)
  u8 *dst = nullptr;       // whole mode: the final stream; segment bytes after the first go straight to their place
  size_t dst_cap = 0;
};

// all shards of one lane, in stream order
static int lane_run(PoolRun &R, Lane *L, std::vector<ShardJob> jobs, std::vector<ShardOut *> outs) {
  Ctx *c = L->c;
  CK(cudaSetDevice(c->device));
  c->cap_override = R.pool->cap_override;
  c->batch_override = R.pool->batch_override;
  const size_t nj = jobs.size();
  const u8 *d_in[2] = {nullptr, nullptr};
  std::future<int> up;
  if (nj) {
    const ShardJob j0 = jobs[0];
    up = std::async(std::launch::async, [L, j0, &R, &d_in] { numa_bind_self(L->c->device); return lane_upload(L, j0, 0, R.pageable, &d_in[0]); });
  }
  bz2b200_stats agg{};
  static const bool ptrace = getenv("BZ2B200_POOL_TRACE") != nullptr;  // development aid: host timeline of every shard
  auto now_ms = [] { return (double)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count() / 1000.0; };
  const double t_origin = now_ms();
  for (size_t k = 0; k < nj; k++) {
    ShardJob &job = jobs[k];
    ShardOut &out = *outs[k];
    const int slot = (int)(k & 1);
    double tp[8] = {now_ms() - t_origin, 0, 0, 0, 0, 0, 0, 0};
    int rc = up.get();
    if (rc) return rc;
    tp[1] = now_ms() - t_origin;
    u64 start = 0;
    bool have_start = job.index == 0;
    for (;;) {  // (again with a longer halo when the last owned block runs out of input)
      if (!job.on_device) CK(cudaStreamWaitEvent(c->stream, L->in_ev[slot], 0));
      if ((rc = pipe_begin(c, d_in[slot], job.n_avail, R.level))) return rc;
      tp[2] = now_ms() - t_origin;
      if (!have_start) { if ((rc = R.ex->get_cut(job.index - 1, &start))) return rc; have_start = true; }
      tp[3] = now_ms() - t_origin;
      const u64 s_local = start > job.base ? start - job.base : 0;
      const i64 own = (i64)(job.own_len < job.n_avail ? job.own_len : job.n_avail);
      if ((rc = pipe_cut(c, (i64)s_local, own))) return rc;
      const auto &h = c->pipe.hrecs;
      const bool at_end = job.last && job.n_avail == job.n_max;
      const bool incomplete = !at_end && !h.empty() && h.back().p == c->pipe.N && h.back().n < c->pipe.B;
      if (!incomplete) break;
      if (job.n_avail >= job.n_max) { c->err = "shard halo too short: the last owned block needs input beyond the buffer"; return BZ2B200_E_ARG; }
      size_t halo = job.n_avail - job.own_len;
      size_t grown = job.own_len + (halo < (1u << 20) ? (8u << 20) : halo * 8);
      job.n_avail = grown < job.n_max ? grown : job.n_max;
      if (job.on_device) continue;
      if ((rc = lane_upload(L, job, slot, R.pageable, &d_in[slot]))) return rc;
    }
    {
      const auto &h = c->pipe.hrecs;
      out.info = bz2b200_shard_info{};
      out.info.n_blocks = (uint32_t)h.size();
      const u64 nxt = h.empty() ? start : job.base + (u64)h.back().p;
      out.info.next_start = nxt > start ? nxt : start;
      out.info.complete = 1;
      if ((rc = R.ex->put_cut(job.index, out.info.next_start))) return rc;
      tp[4] = now_ms() - t_origin;
    }
    if (k + 1 < nj) {  // the next shard of this lane travels while this one is compressed
      const ShardJob jn = jobs[k + 1];
      const int ns = slot ^ 1;
      up = std::async(std::launch::async, [L, jn, ns, &R, &d_in] { numa_bind_self(L->c->device); return lane_upload(L, jn, ns, R.pageable, &d_in[ns]); });
    }
    size_t olen = 0;
    u64 bits = 0;
    u32 fold = 0;
    if ((rc = pipe_run(c, 0, false, nullptr, 0, true, &olen, &bits, &fold))) return rc;
    tp[5] = now_ms() - t_origin;
    trace_report(c);
    {
      agg.n_blocks += c->st.n_blocks; agg.kernel_launches += c->st.kernel_launches; agg.rle1_bytes += c->st.rle1_bytes;
      agg.mtf_syms += c->st.mtf_syms; agg.sort_slots += c->st.sort_slots; agg.d1_triggered |= c->st.d1_triggered;
      if (c->st.sort_rounds > agg.sort_rounds) agg.sort_rounds = c->st.sort_rounds;
      agg.ms_total += c->st.ms_total;
      for (int i = 0; i < 8; i++) agg.ms_stage[i] += c->st.ms_stage[i];
      agg.dom_ms += c->st.dom_ms; agg.dom_launches += c->st.dom_launches; agg.dom_bytes += c->st.dom_bytes;
    }
    out.info.bits = bits;
    out.info.crc_fold = fold;
    ExTail prev;
    if (job.index > 0 && (rc = R.ex->get_tail(job.index - 1, &prev))) return rc;
    tp[6] = now_ms() - t_origin;
    ExTail mine;
    const u32 m = out.info.n_blocks & 31u;
    mine.end_bit = prev.end_bit + bits;
    mine.blocks = prev.blocks + out.info.n_blocks;
    mine.fold = (m ? ((prev.fold << m) | (prev.fold >> (32 - m))) : prev.fold) ^ fold;  // BJ:2237 over this shard's blocks
    if ((rc = R.ex->put_tail(job.index, mine))) return rc;
    out.tail = mine;
    out.bit_off = prev.end_bit;
    // ---- emit: the segment moves to its bit phase and leaves for the host ----
    const u32 phase = (u32)(prev.end_bit & 7);
    out.info.bit_phase = phase;
    const u64 nsrc = (bits + 7) / 8;
    const size_t seg_len = (size_t)((phase + bits + 7) / 8);
    out.seg_bytes = seg_len;
    const u8 *d_seg = P<u8>(c->out);
    if (phase && bits) {
      ENS(c->out2, nsrc + 16);
      LAUNCH(k_shift_bytes, (unsigned)((nsrc + 256) / 256), 256, 0, P<u8>(c->out), nsrc, phase, P<u8>(c->out2));
      d_seg = P<u8>(c->out2);
    }
    if (seg_len && !R.keep_on_device) {
      if (R.dst) {
        const size_t off = (size_t)(prev.end_bit >> 3);
        if (off + seg_len > R.dst_cap) { c->err = "output buffer too small"; return BZ2B200_E_UNEXPECTED_OUTPUT_EOF; }
        RC(rb_add(c, &out.first_byte, d_seg, 1));  // OR-merged by the caller: its byte is shared with the previous segment
        RC(rb_sync(c));
        if (seg_len > 1) CK(cudaMemcpyAsync(R.dst + off + 1, d_seg + 1, seg_len - 1, cudaMemcpyDeviceToHost, c->stream));
      } else {
        out.seg = (u8 *)result_pool().get(seg_len);
        if (!out.seg) return BZ2B200_E_OUT_OF_MEMORY;
        CK(cudaMemcpyAsync(out.seg, d_seg, seg_len, cudaMemcpyDeviceToHost, c->stream));
      }
    }
    tp[7] = now_ms() - t_origin;
    if (ptrace)
      fprintf(stderr, "[bz2b200 pool] dev %d shard %d (%zu MB): start %.2f | upload ready %.2f | begin issued %.2f | cut arrived %.2f | cut done %.2f | stages done %.2f | tail arrived %.2f | emit issued %.2f ms\n",
              c->device, job.index, job.own_len / 1000000, tp[0], tp[1], tp[2], tp[3], tp[4], tp[5], tp[6], tp[7]);
  }
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaGetLastError());
  c->st = agg;
  return 0;
}

// jobs: this process's shards in stream order; shard i goes to lane i mod lanes
static int pool_run_shards(Pool *p, Exchange *ex, std::vector<ShardJob> &jobs, std::vector<ShardOut> &outs, int level, bool pageable, bool keep_on_device,
                           u8 *dst, size_t dst_cap) {
  const size_t nl = p->lanes.size();
  outs.assign(jobs.size(), ShardOut());
  PoolRun R{p, ex, level, pageable || p->force_staging, keep_on_device, dst, dst_cap};
  std::vector<std::vector<ShardJob>> lj(nl);
  std::vector<std::vector<ShardOut *>> lo(nl);
  for (size_t i = 0; i < jobs.size(); i++) { lj[i % nl].push_back(jobs[i]); lo[i % nl].push_back(&outs[i]); }
  std::vector<int> rcs(nl, 0);
  std::vector<std::thread> th;
  for (size_t l = 1; l < nl; l++)
    if (!lj[l].empty()) th.emplace_back([&, l] { numa_bind_self(p->lanes[l]->c->device); rcs[l] = lane_run(R, p->lanes[l], lj[l], lo[l]); if (rcs[l]) ex->fail(rcs[l]); });
  rcs[0] = lane_run(R, p->lanes[0], lj[0], lo[0]);
  if (rcs[0]) ex->fail(rcs[0]);
  for (auto &t : th) t.join();
  p->st = bz2b200_stats{};
  int rc = 0;
  for (size_t l = 0; l < nl; l++) {
    if (rcs[l] && (!rc || rc == BZ2B200_E_PEER)) { rc = rcs[l]; p->err = p->lanes[l]->c->err; }  // the first real failure, not its echo
    if (lj[l].empty()) continue;
    const bz2b200_stats &s = p->lanes[l]->c->st;
    p->st.n_blocks += s.n_blocks; p->st.kernel_launches += s.kernel_launches; p->st.rle1_bytes += s.rle1_bytes; p->st.mtf_syms += s.mtf_syms;
    p->st.sort_slots += s.sort_slots; p->st.d1_triggered |= s.d1_triggered;
    if (s.sort_rounds > p->st.sort_rounds) p->st.sort_rounds = s.sort_rounds;
    p->st.dom_ms += s.dom_ms; p->st.dom_launches += s.dom_launches; p->st.dom_bytes += s.dom_bytes;
    for (int i = 0; i < 8; i++) p->st.ms_stage[i] += s.ms_stage[i];
  }
  return rc;
}

// Shard sizes of one stream.  Small batches of blocks use the GPU less well than large ones (latency-bound per-block
// kernels, partial waves), so the stream is NOT cut evenly: the first shard of every lane is small -- its upload is the
// only one nobody hides -- and every following wave is `growth` times larger (copies run ~2.5x faster than the kernels,
// so the upload of wave k+1 still fits under the kernels of wave k).  `fixed` > 0: equal shards of that size.
static std::vector<size_t> pool_plan(size_t n, int level, size_t lanes, size_t fixed, size_t first, double growth) {
  std::vector<size_t> sizes;
  if (n == 0) { sizes.push_back(0); return sizes; }
  if (fixed) {
    for (size_t o = 0; o < n; o += fixed) sizes.push_back(n - o < fixed ? n - o : fixed);
    return sizes;
  }
  if (growth >= 1000.0 && first && first < n) {  // two shards: `first`, then everything else
    sizes.push_back(first);
    sizes.push_back(n - first);
    return sizes;
  }
  const size_t B = (size_t)level * 100000;
  // Measured on B200 (profiles/r02_pool_plans.md): a batch of few blocks costs ~1 ms more than its share of a large one
  // (per-block CTAs, host round trips), so few shards win: a quarter of a lane's share first (its upload is the exposed
  // one, the rest travels under its kernels), then everything else.
  if (!first) {
    first = n / (4 * lanes) + 1;
    if (first < 6 * B) first = 6 * B;
  }
  if (growth < 1.0) growth = 3.0;
  size_t left = n;
  const double wmax = (double)((size_t)128 << 20);  // bounds the per-lane state (52 B per input byte)
  double w = (double)first < wmax ? (double)first : wmax;
  while (left) {
    if ((double)left <= w * (double)lanes * 1.3) {  // the last wave: what is left, split evenly over the lanes (no crumb shard)
      size_t parts = (double)left <= w * 0.65 ? 1 : (size_t)(((double)left + w * 1.3 - 1) / (w * 1.3));
      if (parts < 1) parts = 1;  // a first shard of one or two bytes: the quotient above can round below 1
      for (size_t l = 0; l < parts; l++) {
        size_t take = l + 1 == parts ? left : ((left / (parts - l)) + 4095) & ~(size_t)4095;
        if (take > left) take = left;
        if (!take) break;  // the rounding gave the earlier parts everything
        sizes.push_back(take);
        left -= take;
      }
      break;
    }
    for (size_t l = 0; l < lanes; l++) {              // a full wave: one shard per lane
      size_t take = ((size_t)w + 4095) & ~(size_t)4095;
      if (take > left) take = left;  // shards far below 4 KiB (tests): the rounding must not run past the input
      if (!take) break;
      sizes.push_back(take);
      left -= take;
    }
    w *= growth;
    if (w > wmax) w = wmax;
  }
  return sizes;
}
static size_t pool_first_halo(const Pool *p, int level) {
  if (p->halo0) return p->halo0;
  const size_t B = p->cap_override ? p->cap_override : (size_t)level * 100000;
  return B + B / 4 + 65536;  // a block of text reads about B bytes; run-heavy input grows the halo on demand (up to 51 x B)
}

// Bzip2.compressFile over every lane of the pool (host input, host output in page-locked result memory)
static int pool_compress_whole(Pool *p, const u8 *in, size_t n, int level, size_t shard_bytes, uint8_t **out, size_t *out_len) {
  if (level < 1 || level > 9) return BZ2B200_E_LEVEL;
  p->err.clear();
  auto t0 = std::chrono::steady_clock::now();
  const std::vector<size_t> plan = pool_plan(n, level, p->lanes.size(), shard_bytes, p->plan_first, p->plan_growth);
  const size_t ns = plan.size();
  const size_t halo = pool_first_halo(p, level);
  std::vector<ShardJob> jobs(ns);
  u64 at = 0;
  for (size_t j = 0; j < ns; j++) {
    ShardJob &J = jobs[j];
    J.base = at;
    at += plan[j];
    J.src = in + J.base;
    J.n_max = n - (size_t)J.base;
    J.own_len = plan[j];
    J.n_avail = J.own_len + halo < J.n_max ? J.own_len + halo : J.n_max;
    J.index = (int)j;
    J.last = true;  // the buffer runs to the end of the input for every shard: a block that reaches n_max ends the stream
  }
  const size_t cap = bz2b200_compress_bound(n, level);
  u8 *dst = (u8 *)result_pool().get(cap);
  if (!dst) return BZ2B200_E_OUT_OF_MEMORY;
  LocalExchange ex((int)ns);
  std::vector<ShardOut> outs;
  int rc = pool_run_shards(p, &ex, jobs, outs, level, n ? is_pageable(in) : false, false, dst, cap);
  if (rc) { result_pool().put(dst); return rc; }
  // assembly: header, the byte each segment shares with its predecessor, footer (BJ:2223-2226, 2245-2247)
  dst[0] = 'B'; dst[1] = 'Z'; dst[2] = 'h'; dst[3] = (u8)('0' + level);
  for (size_t j = 0; j < ns; j++) {
    if (!outs[j].seg_bytes) continue;
    const size_t off = (size_t)(outs[j].bit_off >> 3);
    if (outs[j].bit_off & 7) dst[off] |= outs[j].first_byte; else dst[off] = outs[j].first_byte;
  }
  const ExTail &T = outs[ns - 1].tail;
  u64 bitpos = T.end_bit;
  if ((bitpos & 7) == 0) dst[bitpos >> 3] = 0;
  for (size_t i = (size_t)(bitpos >> 3) + 1; i < (size_t)((bitpos + 80 + 7) >> 3); i++) dst[i] = 0;
  const u64 vals[2] = {BZ_MAGIC_END, (u64)T.fold};
  const int lens[2] = {48, 32};
  for (int q = 0; q < 2; q++)
    for (int i = lens[q] - 1; i >= 0; i--, bitpos++)
      if ((vals[q] >> i) & 1) dst[bitpos >> 3] |= (u8)(0x80u >> (bitpos & 7));
  *out = dst;
  *out_len = (size_t)((bitpos + 7) / 8);
  p->st.in_bytes = n;
  p->st.out_bytes = *out_len;
  p->st.ms_total = (float)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() / 1000.f;
  return BZ2B200_OK;
}

// this process's shards of a stream that spans the ranks of a group (or, with grp == nullptr, of this process alone)
static int pool_compress_ranked(Pool *p, Group *grp, const bz2b200_shard_job *jobs_in, int n_jobs, int total_shards, int level, int keep_on_device,
                                bz2b200_shard_result *results) {
  if (level < 1 || level > 9) return BZ2B200_E_LEVEL;
  if (n_jobs < 0 || total_shards < 1 || total_shards > GRP_SLOTS || (n_jobs && (!jobs_in || !results))) return BZ2B200_E_ARG;
  p->err.clear();
  auto t0 = std::chrono::steady_clock::now();
  const size_t halo = pool_first_halo(p, level);
  std::vector<ShardJob> jobs((size_t)n_jobs);
  bool pageable = false;
  for (int i = 0; i < n_jobs; i++) {
    const bz2b200_shard_job &a = jobs_in[i];
    if (a.index < 0 || a.index >= total_shards || (i && a.index <= jobs_in[i - 1].index) || a.own_len > a.n_readable || (a.n_readable && !a.src) ||
        (a.on_device && ((uintptr_t)a.src & 15)))
      return BZ2B200_E_ARG;
    ShardJob &J = jobs[(size_t)i];
    J.src = (const u8 *)a.src; J.n_max = a.n_readable; J.own_len = a.own_len; J.base = a.base; J.index = a.index; J.on_device = a.on_device != 0;
    J.n_avail = J.own_len + halo < J.n_max ? J.own_len + halo : J.n_max;
    J.last = a.index == total_shards - 1;
    if (!J.on_device && J.n_max && !i) pageable = is_pageable(J.src);
  }
  std::vector<ShardOut> outs;
  int rc;
  if (grp) {
    grp->epoch++;
    GroupExchange ex(grp);
    rc = pool_run_shards(p, &ex, jobs, outs, level, pageable, keep_on_device != 0, nullptr, 0);
    int rb = ex.barrier();
    if (!rc) rc = rb;
  } else {
    LocalExchange ex(total_shards);
    rc = pool_run_shards(p, &ex, jobs, outs, level, pageable, keep_on_device != 0, nullptr, 0);
  }
  for (int i = 0; i < n_jobs && i < (int)outs.size(); i++) {
    if (rc) { if (outs[(size_t)i].seg) result_pool().put(outs[(size_t)i].seg); results[i] = bz2b200_shard_result{}; continue; }
    const ShardOut &o = outs[(size_t)i];
    results[i].seg = o.seg; results[i].seg_bytes = o.seg_bytes; results[i].info = o.info; results[i].bit_offset = o.bit_off;
    results[i].end_bit = o.tail.end_bit; results[i].blocks_through = o.tail.blocks; results[i].crc_fold_through = o.tail.fold;
  }
  p->st.ms_total = (float)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() / 1000.f;
  return rc;
}

// ------------------------------------------------------------------------------------------------ decompress
// Block-range decompression of ONE stream (last paragraph; BJ:1769-1796 walking blocks found by BJ:1434-1439):
// the stream is cut into byte slices; a lane scans its slice for signatures and decodes every candidate at once, then
// receives the state of the reference's walk from the slice before (where the next block starts, the running stream
// CRC, the level), follows it through its own candidates and passes it on -- a chain of scalars like the cut offsets
// of the compressor.  Output offsets are the running sum of decoded sizes (the `tail` chain).
struct DecOut {
  u8 *part = nullptr;     // ranks mode or overflow of the guessed output size: page-locked bytes of this range
  u64 bytes = 0, out_off = 0;
  int rc = 0;
  std::string msg;
  u32 blocks = 0;
};
static inline u64 walk_pack(const DecWalk &w) { return (u64)w.stream_crc | ((u64)w.level << 32) | ((u64)w.ended << 40); }
static inline void walk_unpack(u64 a, u64 b, DecWalk &w) { w.cur = a; w.stream_crc = (u32)b; w.level = (u32)((b >> 32) & 0xff); w.ended = (u32)((b >> 40) & 1); }

struct DecRun {
  Pool *pool;
  Exchange *ex;
  u64 total_n;
  int multistream, first_level;
  bool pageable;
  u8 *dst = nullptr;      // whole mode: the output buffer (capacity guessed from the hint); parts that do not fit stay apart
  size_t dst_cap = 0;
  bool keep_on_device = false;  // ranks mode: the decoded bytes stay in the lane's output slot
};

static int lane_decode(DecRun &R, Lane *L, std::vector<ShardJob> jobs, std::vector<DecOut *> outs) {
  Ctx *c = L->c;
  CK(cudaSetDevice(c->device));
  if (R.pool->dec_batch) c->dec_batch = R.pool->dec_batch;
  const size_t nj = jobs.size();
  const u8 *d_in[2] = {nullptr, nullptr};
  std::future<int> up;
  if (nj) {
    const ShardJob j0 = jobs[0];
    up = std::async(std::launch::async, [L, j0, &R, &d_in] { numa_bind_self(L->c->device); return lane_upload(L, j0, 0, R.pageable, &d_in[0]); });
  }
  bz2b200_stats agg{};
  for (size_t k = 0; k < nj; k++) {
    ShardJob &job = jobs[k];
    DecOut &out = *outs[k];
    const int slot = (int)(k & 1);
    int rc = up.get();
    if (rc) return rc;
    if (k + 1 < nj) {
      const ShardJob jn = jobs[k + 1];
      const int ns = slot ^ 1;
      up = std::async(std::launch::async, [L, jn, ns, &R, &d_in] { numa_bind_self(L->c->device); return lane_upload(L, jn, ns, R.pageable, &d_in[ns]); });
    }
    DecWalk Win, Wout;
    bool have_in = false;
    auto get_walk = [&](DecWalk &w) -> int {
      if (!have_in) {
        if (job.index == 0) { Win = DecWalk(); Win.level = (u32)R.first_level; }
        else { u64 a = 0, b = 0; int r = R.ex->get_cut(job.index - 1, &a, &b); if (r) return r; walk_unpack(a, b, Win); }
        have_in = true;
      }
      w = Win;
      return 0;
    };
    bool cut_sent = false;
    const std::function<int(const DecWalk &)> on_walk = [&](const DecWalk &w) -> int {
      cut_sent = true;
      return R.ex->put_cut(job.index, w.cur, walk_pack(w));
    };
    DecodeResult res;
    const u64 lo_bit = job.base * 8, hi_bit = (job.base + job.own_len) * 8;
    c->st = bz2b200_stats{};
    c->err.clear();
    bool bad_header = false;
    if (job.index == 0) {  // whoever holds the first slice checks the stream header (BJ:1408-1427); the level travels with the walk
      u8 hdr[4] = {0, 0, 0, 0};
      const size_t hb = job.n_avail < 4 ? job.n_avail : 4;
      if (job.on_device) { if (hb) CK(cudaMemcpy(hdr, job.src, hb, cudaMemcpyDeviceToHost)); }
      else memcpy(hdr, job.src, hb);
      bad_header = hb < 4 || hdr[0] != 'B' || hdr[1] != 'Z' || hdr[2] != 'h' || hdr[3] != (u8)('0' + R.first_level);
    }
    for (;;) {
      if (bad_header) { rc = BZ2B200_E_NOT_BZIP_DATA; c->err = "no bzip2 stream header, or not the level the caller announced"; get_walk(Win); break; }
      if (!job.on_device) CK(cudaStreamWaitEvent(c->stream, L->in_ev[slot], 0));
      BufSink sink(&L->out_slot[slot]);
      int need_more = 0;
      rc = decode_range(c, d_in[slot], job.n_avail, job.base, R.total_n, lo_bit, hi_bit, R.multistream, DEC_STREAM, get_walk, sink, res, Wout, &need_more,
                        &on_walk);
      if (rc == BZ2B200_E_CUDA || rc == BZ2B200_E_PEER || rc == BZ2B200_E_ARG) return rc;  // infrastructure: everybody stops
      if (!need_more) break;
      if (job.n_avail >= job.n_max) { c->err = "internal: range needs input beyond the stream"; return BZ2B200_E_ARG; }
      const size_t halo = job.n_avail - job.own_len;
      const size_t grown = job.own_len + (halo < (4u << 20) ? (16u << 20) : halo * 4);
      job.n_avail = grown < job.n_max ? grown : job.n_max;
      if (!job.on_device && (rc = lane_upload(L, job, slot, R.pageable, &d_in[slot]))) return rc;
    }
    trace_report(c);
    // a data error ends the walk here; the slices after it decode nothing, the slices before it still count: the
    // caller reports the error of the lowest slice (the first in stream order, like the reference)
    out.rc = rc;
    out.msg = c->err;
    if (rc) { Wout = Win; Wout.ended = 1; res.out_len = 0; }
    if (!cut_sent && (rc = R.ex->put_cut(job.index, Wout.cur, walk_pack(Wout)))) return rc;
    ExTail prev;
    prev.end_bit = 0;
    if (job.index > 0 && (rc = R.ex->get_tail(job.index - 1, &prev))) return rc;
    ExTail mine;
    mine.end_bit = prev.end_bit + res.out_len;   // `end_bit` carries the running OUTPUT offset here
    mine.blocks = prev.blocks + c->st.n_blocks;
    if ((rc = R.ex->put_tail(job.index, mine))) return rc;
    out.bytes = res.out_len;
    out.out_off = prev.end_bit;
    out.blocks = c->st.n_blocks;
    agg.n_blocks += c->st.n_blocks; agg.kernel_launches += c->st.kernel_launches; agg.rle1_bytes += c->st.rle1_bytes;
    if (out.bytes && !R.keep_on_device) {
      if (R.dst && out.out_off + out.bytes <= R.dst_cap) {
        CK(cudaMemcpyAsync(R.dst + out.out_off, L->out_slot[slot].p, (size_t)out.bytes, cudaMemcpyDeviceToHost, c->stream));
      } else {
        out.part = (u8 *)result_pool().get((size_t)out.bytes);
        if (!out.part) return BZ2B200_E_OUT_OF_MEMORY;
        CK(cudaMemcpyAsync(out.part, L->out_slot[slot].p, (size_t)out.bytes, cudaMemcpyDeviceToHost, c->stream));
      }
    }
  }
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaGetLastError());
  c->st = agg;
  return 0;
}

static int pool_run_decode(Pool *p, DecRun &R, std::vector<ShardJob> &jobs, std::vector<DecOut> &outs) {
  const size_t nl = p->lanes.size();
  outs.assign(jobs.size(), DecOut());
  std::vector<std::vector<ShardJob>> lj(nl);
  std::vector<std::vector<DecOut *>> lo(nl);
  for (size_t i = 0; i < jobs.size(); i++) { lj[i % nl].push_back(jobs[i]); lo[i % nl].push_back(&outs[i]); }
  std::vector<int> rcs(nl, 0);
  std::vector<std::thread> th;
  for (size_t l = 1; l < nl; l++)
    if (!lj[l].empty()) th.emplace_back([&, l] { numa_bind_self(p->lanes[l]->c->device); rcs[l] = lane_decode(R, p->lanes[l], lj[l], lo[l]); if (rcs[l]) R.ex->fail(rcs[l]); });
  rcs[0] = lane_decode(R, p->lanes[0], lj[0], lo[0]);
  if (rcs[0]) R.ex->fail(rcs[0]);
  for (auto &t : th) t.join();
  p->st = bz2b200_stats{};
  int rc = 0;
  for (size_t l = 0; l < nl; l++) {
    if (rcs[l] && (!rc || rc == BZ2B200_E_PEER)) { rc = rcs[l]; p->err = p->lanes[l]->c->err; }
    if (lj[l].empty()) continue;
    const bz2b200_stats &s = p->lanes[l]->c->st;
    p->st.n_blocks += s.n_blocks; p->st.kernel_launches += s.kernel_launches; p->st.rle1_bytes += s.rle1_bytes;
  }
  return rc;
}

static size_t pool_dec_slice(const Pool *p, size_t n) {
  size_t per = n / p->lanes.size() + 1;
  const size_t lo = (size_t)24 << 20, hi = (size_t)352 << 20;  // >= ~80 blocks of text per slice (the parse is latency-bound below), <= one batch of DEC_BATCH
  if (per < lo) per = lo;
  if (per > hi) per = hi;
  return per;
}
static const size_t kDecHalo0 = (size_t)1200000;  // a level-9 block of incompressible data is ~1.1 MB of stream; grown on demand

// Bzip2.decompressFile of one stream over every lane of the pool (host input, host output in page-locked result memory)
static int pool_decompress_whole(Pool *p, const u8 *in, size_t n, int multistream, size_t size_hint, size_t slice_bytes, uint8_t **out, size_t *out_len) {
  p->err.clear();
  auto t0 = std::chrono::steady_clock::now();
  if (n < 4 || in[0] != 'B' || in[1] != 'Z' || in[2] != 'h' || in[3] < '1' || in[3] > '9') return BZ2B200_E_NOT_BZIP_DATA;  // BJ:1408-1427
  if (!slice_bytes) slice_bytes = pool_dec_slice(p, n);
  const size_t ns = (n + slice_bytes - 1) / slice_bytes;
  std::vector<ShardJob> jobs(ns);
  for (size_t j = 0; j < ns; j++) {
    ShardJob &J = jobs[j];
    J.base = (u64)j * slice_bytes;
    J.src = in + J.base;
    J.n_max = n - (size_t)J.base;
    J.own_len = J.n_max < slice_bytes ? J.n_max : slice_bytes;
    const size_t h0 = p->halo0 ? p->halo0 : kDecHalo0;
    J.n_avail = J.own_len + h0 < J.n_max ? J.own_len + h0 : J.n_max;
    J.index = (int)j;
  }
  const size_t cap = size_hint ? size_hint : (n * 6 > ((size_t)1 << 20) ? n * 6 : (size_t)1 << 20);
  u8 *dst = (u8 *)result_pool().get(cap);
  if (!dst) return BZ2B200_E_OUT_OF_MEMORY;
  LocalExchange ex((int)ns);
  DecRun R{p, &ex, (u64)n, multistream, in[3] - '0', p->force_staging || is_pageable(in), dst, cap, false};
  std::vector<DecOut> outs;
  int rc = pool_run_decode(p, R, jobs, outs);
  u64 total = 0;
  bool apart = false;
  for (size_t j = 0; j < ns && !rc; j++) {
    if (outs[j].rc) { rc = outs[j].rc; p->err = outs[j].msg; break; }  // the first error in stream order
    total = outs[j].out_off + outs[j].bytes;
    apart |= outs[j].part != nullptr;
  }
  if (!rc && apart) {  // the guess was too small: one exact buffer, everything copied once more (highly compressible data)
    u8 *big = (u8 *)result_pool().get((size_t)total);
    if (!big) rc = BZ2B200_E_OUT_OF_MEMORY;
    for (size_t j = 0; j < ns && !rc; j++) {
      if (!outs[j].bytes) continue;
      memcpy(big + outs[j].out_off, outs[j].part ? outs[j].part : dst + outs[j].out_off, (size_t)outs[j].bytes);
    }
    if (!rc) { result_pool().put(dst); dst = big; }
  }
  for (size_t j = 0; j < outs.size(); j++) if (outs[j].part) result_pool().put(outs[j].part);
  if (rc) { result_pool().put(dst); return rc; }
  *out = dst;
  *out_len = (size_t)total;
  p->st.in_bytes = n;
  p->st.out_bytes = total;
  p->st.ms_total = (float)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() / 1000.f;
  return BZ2B200_OK;
}

// this process's slices of a stream that spans the ranks of a group: results[i] = the decoded bytes of jobs[i] and
// their offset in the output; the status of the whole stream is the first non-zero results[].rc in slice order
static int pool_decompress_ranked(Pool *p, Group *grp, const bz2b200_shard_job *jobs_in, int n_jobs, int total_shards, uint64_t total_n, int first_level,
                                  int multistream, int keep_on_device, bz2b200_range_result *results) {
  if (n_jobs < 0 || total_shards < 1 || total_shards > GRP_SLOTS || (n_jobs && (!jobs_in || !results)) || first_level < 1 || first_level > 9) return BZ2B200_E_ARG;
  p->err.clear();
  auto t0 = std::chrono::steady_clock::now();
  std::vector<ShardJob> jobs((size_t)n_jobs);
  bool pageable = false;
  for (int i = 0; i < n_jobs; i++) {
    const bz2b200_shard_job &a = jobs_in[i];
    if (a.index < 0 || a.index >= total_shards || (i && a.index <= jobs_in[i - 1].index) || a.own_len > a.n_readable || (a.n_readable && !a.src) ||
        (a.on_device && ((uintptr_t)a.src & 15)))
      return BZ2B200_E_ARG;
    ShardJob &J = jobs[(size_t)i];
    J.src = (const u8 *)a.src; J.n_max = a.n_readable; J.own_len = a.own_len; J.base = a.base; J.index = a.index; J.on_device = a.on_device != 0;
    const size_t h0 = p->halo0 ? p->halo0 : kDecHalo0;
    J.n_avail = J.own_len + h0 < J.n_max ? J.own_len + h0 : J.n_max;
    if (!J.on_device && J.n_max && !i) pageable = is_pageable(J.src);
  }
  std::vector<DecOut> outs;
  int rc;
  if (grp) {
    grp->epoch++;
    GroupExchange ex(grp);
    DecRun R{p, &ex, total_n, multistream, first_level, pageable || p->force_staging, nullptr, 0, keep_on_device != 0};
    rc = pool_run_decode(p, R, jobs, outs);
    int rb = ex.barrier();
    if (!rc) rc = rb;
  } else {
    LocalExchange ex(total_shards);
    DecRun R{p, &ex, total_n, multistream, first_level, pageable || p->force_staging, nullptr, 0, keep_on_device != 0};
    rc = pool_run_decode(p, R, jobs, outs);
  }
  for (int i = 0; i < n_jobs && i < (int)outs.size(); i++) {
    DecOut &o = outs[(size_t)i];
    if (rc || keep_on_device) { if (o.part) result_pool().put(o.part); o.part = nullptr; }
    results[i].part = o.part; results[i].bytes = o.bytes; results[i].out_offset = o.out_off; results[i].rc = o.rc; results[i].n_blocks = o.blocks;
  }
  p->st.ms_total = (float)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() / 1000.f;
  return rc;
}
