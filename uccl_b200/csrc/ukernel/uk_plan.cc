#include "uk_plan.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <sstream>

#include "../coll/comm.h"
#include "../common/log.h"

namespace ub {

namespace {

uint64_t ceil_div(uint64_t a, uint64_t b) { return (a + b - 1) / b; }
uint64_t round_up(uint64_t a, uint64_t b) { return ceil_div(a, b) * b; }

// Emits ops in per-lane program order; every op depends on its predecessor on the same lane.
struct Emitter {
  UkPlan& plan;
  std::vector<int> last;  // last op index per lane
  explicit Emitter(UkPlan& p) : plan(p), last(p.nlanes, -1) {}
  int emit(UkPlanOp op) {
    if (last[op.lane] >= 0) op.deps.push_back(last[op.lane]);
    plan.ops.push_back(op);
    last[op.lane] = (int)plan.ops.size() - 1;
    return last[op.lane];
  }
  void send(int lane, int tile, int step, int peer, UkBuf dbuf, uint64_t doff, UkBuf sbuf, uint64_t soff, uint64_t bytes) {
    UkPlanOp o;
    o.kind = UkPlanOp::Send;
    o.lane = lane, o.tile = tile, o.step = step, o.peer = peer;
    o.dst = {dbuf, doff}, o.src = {sbuf, soff}, o.bytes = bytes;
    emit(o);
  }
  void recv(int lane, int tile, int step, int peer, uint64_t bytes) {
    UkPlanOp o;
    o.kind = UkPlanOp::Recv;
    o.lane = lane, o.tile = tile, o.step = step, o.peer = peer, o.bytes = bytes;
    emit(o);
  }
  void copy(int lane, int tile, int step, UkBuf dbuf, uint64_t doff, UkBuf sbuf, uint64_t soff, uint64_t bytes) {
    UkPlanOp o;
    o.kind = UkPlanOp::Copy;
    o.lane = lane, o.tile = tile, o.step = step;
    o.dst = {dbuf, doff}, o.src = {sbuf, soff}, o.bytes = bytes;
    emit(o);
  }
  void reduce(int lane, int tile, int step, UkBuf dbuf, uint64_t doff, UkBuf abuf, uint64_t aoff, UkBuf bbuf, uint64_t boff,
              uint64_t bytes) {
    UkPlanOp o;
    o.kind = UkPlanOp::Reduce;
    o.lane = lane, o.tile = tile, o.step = step;
    o.dst = {dbuf, doff}, o.src = {abuf, aoff}, o.src2 = {bbuf, boff}, o.bytes = bytes;
    emit(o);
  }
};

// chunk c (one per rank) of a `bytes`-long message, cut into tiles
struct Chunks {
  uint64_t bytes, chunk, tile;
  int n;
  Chunks(uint64_t b, int nranks, uint64_t tile_bytes) : bytes(b), tile(tile_bytes), n(nranks) {
    chunk = round_up(ceil_div(b, (uint64_t)nranks), 16);
  }
  uint64_t chunk_len(int c) const {
    const uint64_t lo = (uint64_t)c * chunk;
    return lo >= bytes ? 0 : std::min(chunk, bytes - lo);
  }
  int tiles() const { return (int)std::max<uint64_t>(1, ceil_div(chunk, tile)); }
  uint64_t off(int c, int t) const { return (uint64_t)c * chunk + (uint64_t)t * tile; }
  uint64_t len(int c, int t) const {
    const uint64_t cl = chunk_len(c), lo = (uint64_t)t * tile;
    return lo >= cl ? 0 : std::min(tile, cl - lo);
  }
};

void check_params(const UkPlanParams& p) {
  UB_CHECK(p.nranks >= 1 && p.rank >= 0 && p.rank < p.nranks, "ukernel plan: bad rank %d/%d", p.rank, p.nranks);
  UB_CHECK(p.nlanes >= 1, "ukernel plan: nlanes must be >= 1");
  UB_CHECK(p.tile_bytes >= 16 && p.tile_bytes % 16 == 0, "ukernel plan: tile_bytes must be a positive multiple of 16");
  UB_CHECK(p.elem_size >= 1 && 16 % p.elem_size == 0, "ukernel plan: element size %lu unsupported",
           (unsigned long)p.elem_size);
}

UkPlan make_plan(UkColl coll, UkAlgo algo, uint64_t bytes, const UkPlanParams& p) {
  UkPlan pl;
  pl.coll = coll, pl.algo = algo;
  pl.nranks = p.nranks, pl.rank = p.rank, pl.nlanes = p.nlanes;
  pl.bytes = bytes, pl.tile_bytes = p.tile_bytes;
  pl.scratch_bytes = (coll == UkColl::AllReduce || coll == UkColl::ReduceScatter)
                         ? uk_scratch_bytes(algo, p.nranks, p.nlanes, p.tile_bytes)
                         : 0;
  return pl;
}

// zero-byte Send to / Recv from every peer: "I have entered this operation"
void handshake(Emitter& e, int lane, int n, int r) {
  for (int k = 1; k < n; ++k) e.send(lane, -1, 0, (r + k) % n, UkBuf::Out, 0, UkBuf::In, 0, 0);
  for (int k = 1; k < n; ++k) e.recv(lane, -1, 0, (r + k) % n, 0);
}

}  // namespace

UkAlgo uk_select_algo(UkColl coll, int nranks, uint64_t bytes) {
  (void)bytes;
  if (coll != UkColl::AllReduce) return UkAlgo::FullMesh;
  // NVSwitch: every peer is one hop away at full bandwidth, so the 2-step full mesh moves the
  // same bytes as the ring in 2 instead of 2(N-1) dependent steps.
  return nranks <= 1 ? UkAlgo::FullMesh : UkAlgo::FullMesh;
}

uint64_t uk_scratch_bytes(UkAlgo algo, int nranks, int nlanes, uint64_t tile_bytes) {
  if (nranks <= 1) return 0;
  const uint64_t slots = algo == UkAlgo::Ring ? (uint64_t)(nranks - 1) : (uint64_t)nranks;
  return (uint64_t)nlanes * slots * tile_bytes;
}

UkPlan uk_plan_allreduce(uint64_t bytes, const UkPlanParams& p) {
  check_params(p);
  UB_CHECK(bytes % p.elem_size == 0, "ukernel plan: %lu bytes is not a whole number of elements", (unsigned long)bytes);
  UkAlgo algo = p.algo == UkAlgo::Auto ? uk_select_algo(UkColl::AllReduce, p.nranks, bytes) : p.algo;
  UkPlan pl = make_plan(UkColl::AllReduce, algo, bytes, p);
  Emitter e(pl);
  const int n = p.nranks, r = p.rank, L = p.nlanes;
  if (n == 1) {
    for (uint64_t off = 0, t = 0; off < bytes; off += p.tile_bytes, ++t)
      e.copy((int)(t % L), (int)t, 0, UkBuf::Out, off, UkBuf::In, off, std::min(p.tile_bytes, bytes - off));
    return pl;
  }
  const Chunks ch(bytes, n, p.tile_bytes);
  const int T = ch.tiles();
  if (algo == UkAlgo::FullMesh) {
    auto slot = [&](int lane, int src) { return ((uint64_t)lane * n + src) * p.tile_bytes; };
    for (int t = 0; t < T; ++t) {
      const int lane = t % L;
      const uint64_t mine = ch.len(r, t);
      // reduce-scatter: my piece of every peer's chunk goes into that peer's scratch slot[me]
      for (int k = 1; k < n; ++k) {
        const int q = (r + k) % n;
        if (ch.len(q, t)) e.send(lane, t, 0, q, UkBuf::Scratch, slot(lane, r), UkBuf::In, ch.off(q, t), ch.len(q, t));
      }
      if (mine) {
        bool first = true;
        for (int k = 1; k < n; ++k) {
          const int q = (r + k) % n;
          e.recv(lane, t, 0, q, mine);
          e.reduce(lane, t, 0, UkBuf::Out, ch.off(r, t), first ? UkBuf::In : UkBuf::Out, ch.off(r, t), UkBuf::Scratch,
                   slot(lane, q), mine);
          first = false;
        }
        // all-gather: the reduced chunk goes straight into every peer's output
        for (int k = 1; k < n; ++k)
          e.send(lane, t, 1, (r + k) % n, UkBuf::Out, ch.off(r, t), UkBuf::Out, ch.off(r, t), mine);
      }
      for (int k = 1; k < n; ++k) {
        const int q = (r + k) % n;
        if (ch.len(q, t)) e.recv(lane, t, 1, q, ch.len(q, t));
      }
    }
  } else {
    const int next = (r + 1) % n, prev = (r + n - 1) % n;
    auto slot = [&](int lane, int step) { return ((uint64_t)lane * (n - 1) + step) * p.tile_bytes; };
    for (int t = 0; t < T; ++t) {
      const int lane = t % L;
      for (int k = 0; k < n - 1; ++k) {  // reduce-scatter
        const int cs = ((r - k) % n + n) % n, cr = ((r - k - 1) % n + n) % n;
        if (ch.len(cs, t))
          e.send(lane, t, k, next, UkBuf::Scratch, slot(lane, k), k == 0 ? UkBuf::In : UkBuf::Out, ch.off(cs, t),
                 ch.len(cs, t));
        if (ch.len(cr, t)) {
          e.recv(lane, t, k, prev, ch.len(cr, t));
          e.reduce(lane, t, k, UkBuf::Out, ch.off(cr, t), UkBuf::In, ch.off(cr, t), UkBuf::Scratch, slot(lane, k),
                   ch.len(cr, t));
        }
      }
      for (int k = 0; k < n - 1; ++k) {  // all-gather, direct placement into the neighbour's output
        const int cs = ((r + 1 - k) % n + n) % n, cr = ((r - k) % n + n) % n;
        if (ch.len(cs, t)) e.send(lane, t, n - 1 + k, next, UkBuf::Out, ch.off(cs, t), UkBuf::Out, ch.off(cs, t), ch.len(cs, t));
        if (ch.len(cr, t)) e.recv(lane, t, n - 1 + k, prev, ch.len(cr, t));
      }
    }
  }
  return pl;
}

UkPlan uk_plan_alltoall(uint64_t block, const UkPlanParams& p) {
  check_params(p);
  UkPlan pl = make_plan(UkColl::AllToAll, UkAlgo::FullMesh, block, p);
  Emitter e(pl);
  const int n = p.nranks, r = p.rank, L = p.nlanes;
  const int T = (int)std::max<uint64_t>(1, ceil_div(block, p.tile_bytes));
  for (int lane = 0; lane < std::min(L, T); ++lane)
    if (n > 1) handshake(e, lane, n, r);
  for (int t = 0; t < T; ++t) {
    const int lane = t % L;
    const uint64_t lo = (uint64_t)t * p.tile_bytes;
    const uint64_t len = lo >= block ? 0 : std::min(p.tile_bytes, block - lo);
    if (!len) continue;
    for (int k = 1; k < n; ++k) {
      const int q = (r + k) % n;
      e.send(lane, t, 1, q, UkBuf::Out, (uint64_t)r * block + lo, UkBuf::In, (uint64_t)q * block + lo, len);
    }
    e.copy(lane, t, 1, UkBuf::Out, (uint64_t)r * block + lo, UkBuf::In, (uint64_t)r * block + lo, len);
    for (int k = 1; k < n; ++k) e.recv(lane, t, 1, (r + k) % n, len);
  }
  return pl;
}

UkPlan uk_plan_allgather(uint64_t block, const UkPlanParams& p) {
  check_params(p);
  UkPlan pl = make_plan(UkColl::AllGather, UkAlgo::FullMesh, block, p);
  Emitter e(pl);
  const int n = p.nranks, r = p.rank, L = p.nlanes;
  const int T = (int)std::max<uint64_t>(1, ceil_div(block, p.tile_bytes));
  for (int lane = 0; lane < std::min(L, T); ++lane)
    if (n > 1) handshake(e, lane, n, r);
  for (int t = 0; t < T; ++t) {
    const int lane = t % L;
    const uint64_t lo = (uint64_t)t * p.tile_bytes;
    const uint64_t len = lo >= block ? 0 : std::min(p.tile_bytes, block - lo);
    if (!len) continue;
    for (int k = 1; k < n; ++k) e.send(lane, t, 1, (r + k) % n, UkBuf::Out, (uint64_t)r * block + lo, UkBuf::In, lo, len);
    e.copy(lane, t, 1, UkBuf::Out, (uint64_t)r * block + lo, UkBuf::In, lo, len);
    for (int k = 1; k < n; ++k) e.recv(lane, t, 1, (r + k) % n, len);
  }
  return pl;
}

UkPlan uk_plan_reduce_scatter(uint64_t block, const UkPlanParams& p) {
  check_params(p);
  UB_CHECK(block % p.elem_size == 0, "ukernel plan: %lu bytes is not a whole number of elements", (unsigned long)block);
  UkPlan pl = make_plan(UkColl::ReduceScatter, UkAlgo::FullMesh, block, p);
  Emitter e(pl);
  const int n = p.nranks, r = p.rank, L = p.nlanes;
  const int T = (int)std::max<uint64_t>(1, ceil_div(block, p.tile_bytes));
  auto slot = [&](int lane, int src) { return ((uint64_t)lane * n + src) * p.tile_bytes; };
  // unlike AllReduce there is no all-gather phase whose receipt proves that the peers are done with their
  // scratch, so every lane starts with the "I have entered" handshake
  for (int lane = 0; lane < std::min(L, T); ++lane)
    if (n > 1) handshake(e, lane, n, r);
  for (int t = 0; t < T; ++t) {
    const int lane = t % L;
    const uint64_t lo = (uint64_t)t * p.tile_bytes;
    const uint64_t len = lo >= block ? 0 : std::min(p.tile_bytes, block - lo);
    if (!len) continue;
    for (int k = 1; k < n; ++k) {
      const int q = (r + k) % n;
      e.send(lane, t, 1, q, UkBuf::Scratch, slot(lane, r), UkBuf::In, (uint64_t)q * block + lo, len);
    }
    bool first = true;
    for (int k = 1; k < n; ++k) {
      const int q = (r + k) % n;
      e.recv(lane, t, 1, q, len);
      e.reduce(lane, t, 1, UkBuf::Out, lo, first ? UkBuf::In : UkBuf::Out, first ? (uint64_t)r * block + lo : lo,
               UkBuf::Scratch, slot(lane, q), len);
      first = false;
    }
    if (n == 1) e.copy(lane, t, 1, UkBuf::Out, lo, UkBuf::In, lo, len);
    // the next tile of this lane reuses the slots: peers must have consumed this tile's data first
    if (n > 1 && t + L < T) handshake(e, lane, n, r);
  }
  return pl;
}

UkPlan uk_plan_broadcast(uint64_t bytes, int root, const UkPlanParams& p) {
  check_params(p);
  UB_CHECK(root >= 0 && root < p.nranks, "ukernel plan: bad root %d", root);
  UkPlan pl = make_plan(UkColl::Broadcast, UkAlgo::FullMesh, bytes, p);
  pl.root = root;
  Emitter e(pl);
  const int n = p.nranks, r = p.rank, L = p.nlanes;
  const int T = (int)std::max<uint64_t>(1, ceil_div(bytes, p.tile_bytes));
  for (int lane = 0; lane < std::min(L, T); ++lane) {
    if (n == 1) break;
    // the root may only write once a peer has entered the operation: peers signal the root, the root waits
    if (r == root) {
      for (int k = 1; k < n; ++k) e.recv(lane, -1, 0, (r + k) % n, 0);
    } else {
      e.send(lane, -1, 0, root, UkBuf::Out, 0, UkBuf::In, 0, 0);
    }
  }
  for (int t = 0; t < T; ++t) {
    const int lane = t % L;
    const uint64_t lo = (uint64_t)t * p.tile_bytes;
    const uint64_t len = lo >= bytes ? 0 : std::min(p.tile_bytes, bytes - lo);
    if (!len) continue;
    if (r == root) {
      for (int k = 1; k < n; ++k) e.send(lane, t, 1, (r + k) % n, UkBuf::Out, lo, UkBuf::In, lo, len);
      e.copy(lane, t, 1, UkBuf::Out, lo, UkBuf::In, lo, len);
    } else {
      e.recv(lane, t, 1, root, len);
    }
  }
  return pl;
}

UkPlan uk_plan_barrier(const UkPlanParams& p) {
  check_params(p);
  UkPlan pl = make_plan(UkColl::Barrier, UkAlgo::FullMesh, 0, p);
  Emitter e(pl);
  if (p.nranks > 1) handshake(e, 0, p.nranks, p.rank);
  return pl;
}

std::string UkPlan::describe() const {
  static const char* coll_names[] = {"allreduce", "alltoall", "allgather", "barrier", "reduce_scatter", "broadcast"};
  static const char* algo_names[] = {"auto", "ring", "fullmesh"};
  static const char* kind_names[] = {"copy", "reduce", "send", "recv"};
  static const char* buf_names[] = {"in", "out", "scratch"};
  std::ostringstream os;
  os << coll_names[(int)coll] << "/" << algo_names[(int)algo] << " rank " << rank << "/" << nranks << " lanes " << nlanes
     << " bytes " << bytes << " tile " << tile_bytes << " ops " << ops.size() << "\n";
  for (size_t i = 0; i < ops.size(); ++i) {
    const auto& o = ops[i];
    os << "  [" << i << "] lane " << o.lane << " tile " << o.tile << " step " << o.step << " " << kind_names[o.kind];
    if (o.kind == UkPlanOp::Send)
      os << " -> r" << o.peer << ":" << buf_names[(int)o.dst.buf] << "+" << o.dst.off << " from " << buf_names[(int)o.src.buf]
         << "+" << o.src.off;
    else if (o.kind == UkPlanOp::Recv)
      os << " <- r" << o.peer;
    else if (o.kind == UkPlanOp::Copy)
      os << " " << buf_names[(int)o.dst.buf] << "+" << o.dst.off << " = " << buf_names[(int)o.src.buf] << "+" << o.src.off;
    else
      os << " " << buf_names[(int)o.dst.buf] << "+" << o.dst.off << " = " << buf_names[(int)o.src.buf] << "+" << o.src.off
         << " (op) " << buf_names[(int)o.src2.buf] << "+" << o.src2.off;
    os << " bytes " << o.bytes;
    if (!o.deps.empty()) {
      os << " deps";
      for (int d : o.deps) os << " " << d;
    }
    os << "\n";
  }
  return os.str();
}

std::string uk_validate(const std::vector<UkPlan>& plans) {
  const int n = (int)plans.size();
  std::ostringstream err;
  // (src, dst, lane) -> sizes in order
  std::map<std::tuple<int, int, int>, std::vector<uint64_t>> sends, recvs;
  for (int r = 0; r < n; ++r) {
    const UkPlan& p = plans[r];
    if (p.rank != r || p.nranks != n) {
      err << "plan " << r << " has rank " << p.rank << "/" << p.nranks;
      return err.str();
    }
    for (size_t i = 0; i < p.ops.size(); ++i) {
      const auto& o = p.ops[i];
      if (o.lane < 0 || o.lane >= p.nlanes) {
        err << "rank " << r << " op " << i << ": lane " << o.lane << " out of range";
        return err.str();
      }
      for (int d : o.deps)
        if (d < 0 || d >= (int)i) {
          err << "rank " << r << " op " << i << ": dependency " << d << " is not an earlier op";
          return err.str();
        }
      if (o.kind == UkPlanOp::Send || o.kind == UkPlanOp::Recv) {
        if (o.peer < 0 || o.peer >= n || o.peer == r) {
          err << "rank " << r << " op " << i << ": bad peer " << o.peer;
          return err.str();
        }
        if (o.kind == UkPlanOp::Send) sends[{r, o.peer, o.lane}].push_back(o.bytes);
        else recvs[{o.peer, r, o.lane}].push_back(o.bytes);
      }
      if (o.kind == UkPlanOp::Send && o.dst.buf == UkBuf::Scratch && o.dst.off + o.bytes > plans[o.peer].scratch_bytes) {
        err << "rank " << r << " op " << i << ": scratch overflow";
        return err.str();
      }
    }
  }
  for (auto& kv : sends) {
    auto it = recvs.find(kv.first);
    if (it == recvs.end() || it->second != kv.second) {
      err << "sends of rank " << std::get<0>(kv.first) << " to rank " << std::get<1>(kv.first) << " on lane "
          << std::get<2>(kv.first) << " do not match the receiver's recvs";
      return err.str();
    }
  }
  for (auto& kv : recvs)
    if (!sends.count(kv.first)) {
      err << "rank " << std::get<1>(kv.first) << " expects data from rank " << std::get<0>(kv.first) << " on lane "
          << std::get<2>(kv.first) << " that is never sent";
      return err.str();
    }
  return "";
}

std::string uk_check_bounds(const UkPlan& plan, uint64_t in_bytes, uint64_t out_bytes, uint64_t scratch_cap,
                            int max_lanes, uint64_t elem_size) {
  std::ostringstream err;
  if (plan.nlanes < 1 || plan.nlanes > max_lanes) {
    err << "plan uses " << plan.nlanes << " lanes, the communicator has " << max_lanes;
    return err.str();
  }
  if (plan.scratch_bytes > scratch_cap) {
    err << "plan needs " << plan.scratch_bytes << " scratch bytes, the communicator has " << scratch_cap;
    return err.str();
  }
  auto cap = [&](const UkRef& r) -> uint64_t {
    switch (r.buf) {
      case UkBuf::In: return in_bytes;
      case UkBuf::Out: return out_bytes;
      default: return plan.scratch_bytes;
    }
  };
  auto bad = [&](const UkRef& r, uint64_t bytes) {
    return (int)r.buf < 0 || (int)r.buf > 2 || (r.off & 15) != 0 || r.off > cap(r) || bytes > cap(r) - r.off;
  };
  for (size_t i = 0; i < plan.ops.size(); ++i) {
    const UkPlanOp& o = plan.ops[i];
    bool wrong = false;
    switch (o.kind) {
      case UkPlanOp::Copy: wrong = bad(o.dst, o.bytes) || bad(o.src, o.bytes); break;
      case UkPlanOp::Reduce:
        wrong = bad(o.dst, o.bytes) || bad(o.src, o.bytes) || bad(o.src2, o.bytes) || elem_size == 0 ||
                o.bytes % elem_size != 0;
        break;
      case UkPlanOp::Send: wrong = o.bytes != 0 && (bad(o.dst, o.bytes) || bad(o.src, o.bytes)); break;
      case UkPlanOp::Recv: break;
      default: wrong = true;
    }
    if (o.lane < 0 || o.lane >= plan.nlanes) wrong = true;
    if ((o.kind == UkPlanOp::Send || o.kind == UkPlanOp::Recv) && (o.peer < 0 || o.peer >= plan.nranks || o.peer == plan.rank))
      wrong = true;
    if (wrong) {
      err << "rank " << plan.rank << " op " << i << " (kind " << o.kind << ", lane " << o.lane << ", peer " << o.peer << ", "
          << o.bytes << " bytes): outside its buffers, misaligned (16 bytes) or malformed";
      return err.str();
    }
  }
  return "";
}

std::string uk_simulate(const std::vector<UkPlan>& plans, const UkSimBuffers& b, int dtype, int redop) {
  const int n = (int)plans.size();
  auto ptr = [&](int rank, const UkRef& ref) -> char* {
    switch (ref.buf) {
      case UkBuf::In: return b.in[rank] + ref.off;
      case UkBuf::Out: return b.out[rank] + ref.off;
      default: return b.scratch[rank] + ref.off;
    }
  };
  // per (rank, lane) program counters over the ops of that lane
  std::vector<std::vector<std::vector<int>>> prog(n);
  std::vector<std::vector<size_t>> pc(n);
  for (int r = 0; r < n; ++r) {
    prog[r].resize(plans[r].nlanes);
    pc[r].assign(plans[r].nlanes, 0);
    for (size_t i = 0; i < plans[r].ops.size(); ++i) prog[r][plans[r].ops[i].lane].push_back((int)i);
  }
  std::map<std::tuple<int, int, int>, uint64_t> sent, got;  // (src, dst, lane)
  // greedy schedule: run every (rank, lane) until it blocks; a rank far ahead of its peers is
  // exactly the situation in which an unsafe scratch reuse would corrupt data
  bool progress = true;
  while (progress) {
    progress = false;
    for (int r = 0; r < n; ++r)
      for (int l = 0; l < plans[r].nlanes; ++l) {
        while (pc[r][l] < prog[r][l].size()) {
          const UkPlanOp& o = plans[r].ops[prog[r][l][pc[r][l]]];
          if (o.kind == UkPlanOp::Recv) {
            auto key = std::make_tuple(o.peer, r, l);
            if (sent[key] <= got[key]) break;
            ++got[key];
          } else if (o.kind == UkPlanOp::Send) {
            if (o.bytes) memcpy(ptr(o.peer, o.dst), ptr(r, o.src), o.bytes);
            ++sent[std::make_tuple(r, o.peer, l)];
          } else if (o.kind == UkPlanOp::Copy) {
            if (ptr(r, o.dst) != ptr(r, o.src)) memmove(ptr(r, o.dst), ptr(r, o.src), o.bytes);
          } else {
            const void* srcs[2] = {ptr(r, o.src), ptr(r, o.src2)};
            host_reduce_n(ptr(r, o.dst), srcs, 2, o.bytes / dtype_size(dtype), dtype, redop, 1.0f);
          }
          ++pc[r][l];
          progress = true;
        }
      }
  }
  for (int r = 0; r < n; ++r)
    for (int l = 0; l < plans[r].nlanes; ++l)
      if (pc[r][l] < prog[r][l].size()) {
        std::ostringstream os;
        os << "deadlock: rank " << r << " lane " << l << " stuck at op " << prog[r][l][pc[r][l]];
        return os.str();
      }
  return "";
}

}  // namespace ub
