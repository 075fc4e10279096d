// oracle/cpu_engine.cpp — TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.
//
// "Restated reference" CPU engine used as the timed CPU baseline (bench.py cpu_baseline and
// `bench.py --impl reference`).  The reference's real path (nanobind + libucp, UCX 1.18.1)
// cannot be built in this image, so this file restates its EXECUTION SHAPE
// op for op, with oracle/tagmatch.c standing in for libucp's tag matching:
//   * one busy-spinning native worker thread per Client/Server object
//       (reference src/bindings/main.cpp:361-468, 1126-1268),
//   * a 1-slot compare-and-swap mailbox per op kind between the Python thread and the worker,
//     at most one op of each kind posted per loop iteration (reference src/bindings/chan.hpp:84-120),
//   * one heap-allocated future per op, completion delivered by calling a Python callable from the
//     worker thread (GIL acquired per completion; reference main.cpp:172-232, 990-1061),
//   * eager messages complete inline on the sender, large messages complete after the receiver
//     has copied them (rendezvous), copies are single-threaded memcpy on a worker thread
//     (what UCX's intra-node self/shm transports do for host memory),
//   * close cancels pending receives with "Request canceled" (reference main.cpp:484-507).
// Transport: same-process only (the shape of the reference's own tests and pingpong.py,
// where Server and Client live in one process).
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "tagmatch.h"

namespace {

typedef void (*done_cb_t)(void* user, uint64_t sender_tag, uint64_t length);
typedef void (*fail_cb_t)(void* user, const char* reason);

constexpr size_t RNDV_THRESHOLD = 8192;  // UCX default intra-node rendezvous threshold is O(8 KiB)

// 1-slot mailbox: 0 spare, 1 loading, 2 ready, 3 closed (reference chan.hpp:118)
template <class T>
struct Slot {
  std::atomic<uint8_t> st{0};
  T data;
  bool wait_emplace(const T& v) {
    for (;;) {
      uint8_t exp = 0;
      if (st.compare_exchange_weak(exp, 1, std::memory_order_acq_rel)) {
        data = v;
        st.store(2, std::memory_order_release);
        return true;
      }
      if (exp == 3) return false;
      std::this_thread::yield();
    }
  }
  template <class F>
  bool try_consume(F&& f) {
    if (st.load(std::memory_order_acquire) != 2) return false;
    f(data);
    st.store(0, std::memory_order_release);
    return true;
  }
  void close() { st.store(3, std::memory_order_release); }
};

struct Future {
  done_cb_t done;
  fail_cb_t fail;
  void* user;
};
struct SendArgs {
  Future* fut;
  int ep;
  uint64_t tag;
  const uint8_t* ptr;
  size_t len;
};
struct RecvArgs {
  Future* fut;
  uint64_t tag, mask;
  uint8_t* ptr;
  size_t cap;
};
struct FlushArgs {
  Future* fut;
};

struct Worker;
struct WireMsg {   // what travels between two workers
  int kind;        // 0 eager, 1 rts, 2 fin
  int src_ep;      // endpoint index at the receiver
  uint64_t tag;
  uint8_t* eager;  // malloc'ed copy (eager)
  const uint8_t* src;  // sender buffer (rts)
  size_t len;
  Worker* sender;
  Future* send_fut;  // completed by FIN
};

struct Worker {
  orc_worker* tm = orc_worker_new();
  std::thread thr;
  std::atomic<int> status{0};  // 0 void, 2 running, 3 closing, 4 closed
  Slot<SendArgs> send_slot;
  Slot<RecvArgs> recv_slot;
  Slot<FlushArgs> flush_slot;
  std::mutex wire_mu;
  std::deque<WireMsg> wire;  // inbound
  std::vector<std::pair<Worker*, int>> eps;  // peer worker, endpoint index at the peer
  std::map<uint64_t, Future*> recv_futs;      // op id -> future
  std::map<uint64_t, WireMsg> rts_pending;    // rendezvous id -> descriptor (unexpected RTS)
  std::vector<Future*> flush_futs;
  uint64_t next_op = 1, next_rts = 1;
  std::atomic<uint64_t> outstanding_sends{0};
  Future* close_fut = nullptr;

  void push_wire(const WireMsg& m) {
    std::lock_guard<std::mutex> lk(wire_mu);
    wire.push_back(m);
  }

  void complete_recv(const orc_match& m, const WireMsg* rts) {
    auto it = recv_futs.find(m.op_id);
    Future* f = it->second;
    recv_futs.erase(it);
    recv_bufs.erase(m.op_id);
    if (m.status == ORC_OK)
      f->done(f->user, m.sender_tag, m.length);
    else
      f->fail(f->user, "Message truncated");
    delete f;
    (void)rts;
  }

  // the ucp_worker_progress stand-in: drain the wire, match, copy, fire callbacks
  void progress() {
    std::deque<WireMsg> batch;
    {
      std::lock_guard<std::mutex> lk(wire_mu);
      batch.swap(wire);
    }
    for (WireMsg& m : batch) {
      if (m.kind == 2) {  // FIN: the receiver copied our rendezvous payload
        m.send_fut->done(m.send_fut->user, 0, 0);
        delete m.send_fut;
        outstanding_sends.fetch_sub(1);
        continue;
      }
      orc_match out;
      if (m.kind == 0) {
        int hit = orc_arrive(tm, (uint32_t)m.src_ep, m.tag, m.eager, m.len, 0, &out);
        free(m.eager);
        if (hit) complete_recv(out, nullptr);
      } else {
        uint64_t id = next_rts++;
        int hit = orc_arrive(tm, (uint32_t)m.src_ep, m.tag, nullptr, m.len, id, &out);
        if (hit) {
          finish_rndv(out, m);
        } else {
          rts_pending[id] = m;
        }
      }
    }
  }
  std::map<uint64_t, uint8_t*> recv_bufs;  // op id -> destination (for rendezvous copies)
  void finish_rndv(const orc_match& out, const WireMsg& m) {
    if (out.status == ORC_OK) memcpy(recv_bufs[out.op_id], m.src, m.len);  // single-threaded copy
    WireMsg fin{};
    fin.kind = 2;
    fin.send_fut = m.send_fut;
    m.sender->push_wire(fin);
    complete_recv(out, &m);
  }

  void run() {
    while (status.load(std::memory_order_acquire) == 2) {
      progress();
      send_slot.try_consume([&](SendArgs& a) {
        Worker* peer = eps[(size_t)a.ep].first;
        WireMsg m{};
        m.src_ep = eps[(size_t)a.ep].second;
        m.tag = a.tag;
        m.len = a.len;
        m.sender = this;
        if (a.len <= RNDV_THRESHOLD) {
          m.kind = 0;
          m.eager = (uint8_t*)malloc(a.len ? a.len : 1);
          memcpy(m.eager, a.ptr, a.len);
          peer->push_wire(m);
          a.fut->done(a.fut->user, 0, 0);  // inline completion (reference main.cpp:372-378)
          delete a.fut;
        } else {
          m.kind = 1;
          m.src = a.ptr;
          m.send_fut = a.fut;
          outstanding_sends.fetch_add(1);
          peer->push_wire(m);
        }
      });
      recv_slot.try_consume([&](RecvArgs& a) {
        uint64_t op = next_op++;
        recv_futs[op] = a.fut;
        recv_bufs[op] = a.ptr;
        orc_match out;
        if (orc_post_recv(tm, op, a.tag, a.mask, a.ptr, a.cap, &out)) {
          if (out.user) {  // an unexpected rendezvous request
            WireMsg m = rts_pending[out.user];
            rts_pending.erase(out.user);
            finish_rndv(out, m);
          } else {
            complete_recv(out, nullptr);
          }
        }
      });
      flush_slot.try_consume([&](FlushArgs& a) { flush_futs.push_back(a.fut); });
      if (!flush_futs.empty() && outstanding_sends.load() == 0) {
        for (Future* f : flush_futs) {
          f->done(f->user, 0, 0);
          delete f;
        }
        flush_futs.clear();
      }
    }
    // shutdown (reference main.cpp:469-550)
    progress();
    // ops still sitting in a mailbox are cancelled (reference cancel_pending_reqs, main.cpp:680-701)
    send_slot.try_consume([&](SendArgs& a) {
      a.fut->fail(a.fut->user, "Request canceled");
      delete a.fut;
    });
    recv_slot.try_consume([&](RecvArgs& a) {
      a.fut->fail(a.fut->user, "Request canceled");
      delete a.fut;
    });
    flush_slot.try_consume([&](FlushArgs& a) {
      a.fut->fail(a.fut->user, "Request canceled");
      delete a.fut;
    });
    send_slot.close();
    recv_slot.close();
    flush_slot.close();
    std::vector<uint64_t> ids(recv_futs.size() + 1);
    orc_cancel_all(tm, ids.data(), ids.size());
    for (auto& kv : recv_futs) {
      kv.second->fail(kv.second->user, "Request canceled");
      delete kv.second;
    }
    recv_futs.clear();
    for (Future* f : flush_futs) {
      f->fail(f->user, "Request canceled");
      delete f;
    }
    flush_futs.clear();
    status.store(4, std::memory_order_release);
    if (close_fut) {
      close_fut->done(close_fut->user, 0, 0);
      delete close_fut;
      close_fut = nullptr;
    }
  }
  void start() {
    eps.reserve(256);
    status.store(2, std::memory_order_release);
    thr = std::thread([this] { run(); });
  }
};

std::mutex g_mu;
std::map<int, Worker*> g_listeners;  // "port" -> server

}  // namespace

extern "C" {

void* swc_worker_new(void) { return new (std::nothrow) Worker(); }

void swc_worker_free(void* wv) {
  Worker* w = (Worker*)wv;
  if (!w) return;
  int st = w->status.load();
  if (st == 2) w->status.store(3);
  if (w->thr.joinable()) w->thr.join();
  {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto it = g_listeners.begin(); it != g_listeners.end();)
      it = (it->second == w) ? g_listeners.erase(it) : ++it;
  }
  orc_worker_free(w->tm);
  delete w;
}

int swc_listen(void* wv, int port) {
  Worker* w = (Worker*)wv;
  if (w->status.load() != 0) return -1;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_listeners.count(port)) return -2;
  g_listeners[port] = w;
  w->start();
  return 0;
}

// returns the endpoint index on the server side, <0 on failure ("Endpoint is not connected")
int swc_connect(void* wv, int port) {
  Worker* c = (Worker*)wv;
  if (c->status.load() != 0) return -1;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_listeners.find(port);
  if (it == g_listeners.end() || it->second->status.load() != 2) return -24;
  Worker* s = it->second;
  int s_index = (int)s->eps.size();
  s->eps.push_back({c, 0});
  c->eps.push_back({s, s_index});
  c->start();
  return s_index;
}

static Future* mk(done_cb_t d, fail_cb_t f, void* u) {
  Future* fut = new Future();
  fut->done = d;
  fut->fail = f;
  fut->user = u;
  return fut;
}

int swc_send(void* wv, int ep, const void* ptr, size_t len, uint64_t tag, done_cb_t d, fail_cb_t f, void* u) {
  Worker* w = (Worker*)wv;
  if (w->status.load(std::memory_order_acquire) != 2) return -1;
  Future* fut = mk(d, f, u);
  if (!w->send_slot.wait_emplace(SendArgs{fut, ep, tag, (const uint8_t*)ptr, len})) {
    fut->fail(fut->user, "Endpoint is not connected");
    delete fut;
  }
  return 0;
}
int swc_recv(void* wv, void* ptr, size_t cap, uint64_t tag, uint64_t mask, done_cb_t d, fail_cb_t f, void* u) {
  Worker* w = (Worker*)wv;
  if (w->status.load(std::memory_order_acquire) != 2) return -1;
  Future* fut = mk(d, f, u);
  if (!w->recv_slot.wait_emplace(RecvArgs{fut, tag, mask, (uint8_t*)ptr, cap})) {
    fut->fail(fut->user, "Endpoint is not connected");
    delete fut;
  }
  return 0;
}
int swc_flush(void* wv, done_cb_t d, fail_cb_t f, void* u) {
  Worker* w = (Worker*)wv;
  if (w->status.load(std::memory_order_acquire) != 2) return -1;
  Future* fut = mk(d, f, u);
  if (!w->flush_slot.wait_emplace(FlushArgs{fut})) {
    fut->fail(fut->user, "Endpoint is not connected");
    delete fut;
  }
  return 0;
}
int swc_close(void* wv, done_cb_t d, void* u) {
  Worker* w = (Worker*)wv;
  int exp = 2;
  w->close_fut = mk(d, nullptr, u);
  if (!w->status.compare_exchange_strong(exp, 3)) {
    delete w->close_fut;
    w->close_fut = nullptr;
    return -1;
  }
  return 0;
}
int swc_status(void* wv) { return ((Worker*)wv)->status.load(); }

}  // extern "C"
