// Dependency engine: versioned variables with read/write dependency tracking, a priority worker pool,
// WaitForVar / WaitAll / Stop / Start.
//
// Capability parity with the reference's ThreadedEngine (src/engine/threaded_engine.{h,cc}:
// ThreadedVar::AppendRead/WriteDependency :52-88, CompleteRead/WriteDependency :90-192, Push :288-313,
// OnComplete :420-470, WaitForVar :375-410, WaitForAll :412-418; worker pools
// threaded_engine_perdevice.cc:76-198).  Design differs: instead of a lock-free linked list of
// VersionedVarBlocks per variable, each variable keeps a small queue of pending requests under its own
// mutex and a counter of running readers -- the same semantics (concurrent reads, exclusive ordered
// writes, FIFO fairness between a writer and later readers) with far less code, which is adequate because
// in this framework the engine schedules host-side work (IO, checkpoints, host reductions, membership
// RPCs), while device kernels are ordered by CUDA streams.
#include <string.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <queue>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../c_api.h"

namespace mxb {

struct Opr;

struct Var {
  std::mutex mu;
  // pending requests in arrival order: (op, is_write)
  std::deque<std::pair<Opr*, bool>> pending;
  int running_readers = 0;
  bool running_writer = false;
  uint64_t version = 0;
};

struct Opr {
  std::function<void()> fn;
  std::vector<Var*> const_vars;
  std::vector<Var*> mutable_vars;
  std::atomic<int> wait{0};
  int priority = 0;
  uint64_t seq = 0;
  std::string name;          // profiler label (empty: anonymous)
  int64_t push_us = 0;       // wall clock at Push (profiling only)
};

// One executed operation as the profiler sees it (reference: ThreadedEngine::ExecuteOprBlock wraps the run in
// profiler start / stop when profiling is on, src/engine/threaded_engine.h:338-347; OprExecStat in src/profiler).
struct ProfRec {
  std::string name;
  uint64_t seq;
  int64_t push_us, start_us, end_us;
  uint64_t thread;
  int priority;
};

static int64_t wall_us() {
  return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
}

struct OprCmp {
  bool operator()(const Opr* a, const Opr* b) const {
    if (a->priority != b->priority) return a->priority < b->priority;  // higher priority first
    return a->seq > b->seq;                                            // then FIFO
  }
};

class Engine {
 public:
  explicit Engine(int nworkers) : nworkers_(nworkers < 1 ? 1 : nworkers) { Start(); }
  ~Engine() {
    WaitAll();
    Stop();
    for (Var* v : vars_) delete v;
  }

  Var* NewVar() {
    Var* v = new Var();
    std::lock_guard<std::mutex> lk(vars_mu_);
    vars_.push_back(v);
    return v;
  }

  void Push(std::function<void()> fn, std::vector<Var*> cv, std::vector<Var*> mv, int priority, const char* name = nullptr) {
    // de-duplicate: a var that is both read and written counts as written (engine.h NewOperator dedupe)
    std::vector<Var*> cv2;
    for (Var* v : cv) {
      bool dup = false;
      for (Var* m : mv) dup |= (m == v);
      for (Var* c : cv2) dup |= (c == v);
      if (!dup) cv2.push_back(v);
    }
    std::vector<Var*> mv2;
    for (Var* v : mv) {
      bool dup = false;
      for (Var* m : mv2) dup |= (m == v);
      if (!dup) mv2.push_back(v);
    }
    Opr* op = new Opr();
    op->fn = std::move(fn);
    op->const_vars = std::move(cv2);
    op->mutable_vars = std::move(mv2);
    op->priority = priority;
    op->seq = seq_.fetch_add(1);
    if (profiling_.load(std::memory_order_relaxed)) {
      if (name) op->name = name;
      op->push_us = wall_us();
    }
    pending_.fetch_add(1);
    // wait = #deps + 1; the extra one is released at the end so the op cannot fire while being registered
    op->wait.store((int)(op->const_vars.size() + op->mutable_vars.size()) + 1);
    for (Var* v : op->const_vars) AppendRead(v, op);
    for (Var* v : op->mutable_vars) AppendWrite(v, op);
    DecWait(op);
  }

  void WaitForVar(Var* v) {
    std::mutex mu;
    std::condition_variable cv;
    bool done = false;
    Push([&] {
      std::lock_guard<std::mutex> lk(mu);
      done = true;
      cv.notify_all();
    }, {v}, {}, 1 << 20);
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return done; });
  }

  void WaitAll() {
    std::unique_lock<std::mutex> lk(done_mu_);
    done_cv_.wait(lk, [&] { return pending_.load() == 0; });
  }

  void Stop() {
    {
      std::lock_guard<std::mutex> lk(q_mu_);
      if (!running_) return;
      running_ = false;
    }
    q_cv_.notify_all();
    for (auto& t : workers_) t.join();
    workers_.clear();
  }

  void Start() {
    std::lock_guard<std::mutex> lk(q_mu_);
    if (running_) return;
    running_ = true;
    for (int i = 0; i < nworkers_; ++i) workers_.emplace_back([this] { WorkerLoop(); });
  }

  int64_t Pending() const { return pending_.load(); }

  void SetProfiling(bool on) { profiling_.store(on); }
  // tab-separated records, one per line: name seq push_us start_us end_us thread priority
  std::string DrainProfile() {
    std::vector<ProfRec> recs;
    {
      std::lock_guard<std::mutex> lk(prof_mu_);
      recs.swap(prof_);
    }
    std::ostringstream os;
    for (const ProfRec& r : recs)
      os << r.name << '\t' << r.seq << '\t' << r.push_us << '\t' << r.start_us << '\t' << r.end_us << '\t' << r.thread << '\t'
         << r.priority << '\n';
    return os.str();
  }

 private:
  void AppendRead(Var* v, Opr* op) {
    bool ready = false;
    {
      std::lock_guard<std::mutex> lk(v->mu);
      if (v->pending.empty() && !v->running_writer) {
        v->running_readers++;
        ready = true;
      } else {
        v->pending.emplace_back(op, false);
      }
    }
    if (ready) DecWait(op);
  }
  void AppendWrite(Var* v, Opr* op) {
    bool ready = false;
    {
      std::lock_guard<std::mutex> lk(v->mu);
      if (v->pending.empty() && !v->running_writer && v->running_readers == 0) {
        v->running_writer = true;
        ready = true;
      } else {
        v->pending.emplace_back(op, true);
      }
    }
    if (ready) DecWait(op);
  }
  void CompleteRead(Var* v) {
    Opr* next = nullptr;
    {
      std::lock_guard<std::mutex> lk(v->mu);
      v->running_readers--;
      if (v->running_readers == 0 && !v->pending.empty() && v->pending.front().second) {
        next = v->pending.front().first;
        v->pending.pop_front();
        v->running_writer = true;
      }
    }
    if (next) DecWait(next);
  }
  void CompleteWrite(Var* v) {
    std::vector<Opr*> ready;
    {
      std::lock_guard<std::mutex> lk(v->mu);
      v->running_writer = false;
      v->version++;
      if (!v->pending.empty()) {
        if (v->pending.front().second) {
          ready.push_back(v->pending.front().first);
          v->pending.pop_front();
          v->running_writer = true;
        } else {
          while (!v->pending.empty() && !v->pending.front().second) {
            ready.push_back(v->pending.front().first);
            v->pending.pop_front();
            v->running_readers++;
          }
        }
      }
    }
    for (Opr* o : ready) DecWait(o);
  }
  void DecWait(Opr* op) {
    if (op->wait.fetch_sub(1) == 1) {
      {
        std::lock_guard<std::mutex> lk(q_mu_);
        queue_.push(op);
      }
      q_cv_.notify_one();
    }
  }
  void WorkerLoop() {
    for (;;) {
      Opr* op = nullptr;
      {
        std::unique_lock<std::mutex> lk(q_mu_);
        q_cv_.wait(lk, [&] { return !running_ || !queue_.empty(); });
        if (queue_.empty()) {
          if (!running_) return;
          continue;
        }
        op = queue_.top();
        queue_.pop();
      }
      const bool prof = profiling_.load(std::memory_order_relaxed) && op->push_us != 0;
      const int64_t t_start = prof ? wall_us() : 0;
      try {
        op->fn();
      } catch (...) {
        // python trampolines never throw; native ops report through their own channels
      }
      if (prof) {
        ProfRec r{op->name, op->seq, op->push_us, t_start, wall_us(),
                  (uint64_t)std::hash<std::thread::id>()(std::this_thread::get_id()) % 100000, op->priority};
        std::lock_guard<std::mutex> lk(prof_mu_);
        if (prof_.size() < (size_t)1 << 20) prof_.push_back(std::move(r));
      }
      for (Var* v : op->const_vars) CompleteRead(v);
      for (Var* v : op->mutable_vars) CompleteWrite(v);
      delete op;
      if (pending_.fetch_sub(1) == 1) {
        std::lock_guard<std::mutex> lk(done_mu_);
        done_cv_.notify_all();
      }
    }
  }

  int nworkers_;
  bool running_ = false;
  std::vector<std::thread> workers_;
  std::priority_queue<Opr*, std::vector<Opr*>, OprCmp> queue_;
  std::mutex q_mu_;
  std::condition_variable q_cv_;
  std::atomic<int64_t> pending_{0};
  std::atomic<uint64_t> seq_{0};
  std::mutex done_mu_;
  std::condition_variable done_cv_;
  std::mutex vars_mu_;
  std::vector<Var*> vars_;
  std::atomic<bool> profiling_{false};
  std::mutex prof_mu_;
  std::vector<ProfRec> prof_;
};

}  // namespace mxb

using mxb::Engine;
using mxb::Var;

typedef void (*MXBEngineFn)(void*);

MXB_API int MXBEngineCreate(int nworkers, void** out) {
  MXB_TRY
  *out = new Engine(nworkers);
  MXB_CATCH
}
MXB_API int MXBEngineFree(void* h) {
  MXB_TRY
  delete reinterpret_cast<Engine*>(h);
  MXB_CATCH
}
MXB_API int MXBEngineNewVar(void* h, void** out) {
  MXB_TRY
  *out = reinterpret_cast<Engine*>(h)->NewVar();
  MXB_CATCH
}
MXB_API int MXBEnginePush(void* h, MXBEngineFn fn, void* arg, void** cvars, int ncv, void** mvars, int nmv,
                          int priority) {
  MXB_TRY
  std::vector<Var*> cv, mv;
  for (int i = 0; i < ncv; ++i) cv.push_back(reinterpret_cast<Var*>(cvars[i]));
  for (int i = 0; i < nmv; ++i) mv.push_back(reinterpret_cast<Var*>(mvars[i]));
  reinterpret_cast<Engine*>(h)->Push([fn, arg] { fn(arg); }, std::move(cv), std::move(mv), priority);
  MXB_CATCH
}
MXB_API int MXBEnginePushNamed(void* h, MXBEngineFn fn, void* arg, void** cvars, int ncv, void** mvars, int nmv, int priority,
                               const char* name) {
  MXB_TRY
  std::vector<Var*> cv, mv;
  for (int i = 0; i < ncv; ++i) cv.push_back(reinterpret_cast<Var*>(cvars[i]));
  for (int i = 0; i < nmv; ++i) mv.push_back(reinterpret_cast<Var*>(mvars[i]));
  reinterpret_cast<Engine*>(h)->Push([fn, arg] { fn(arg); }, std::move(cv), std::move(mv), priority, name);
  MXB_CATCH
}
MXB_API int MXBEngineSetProfiling(void* h, int on) {
  MXB_TRY
  reinterpret_cast<Engine*>(h)->SetProfiling(on != 0);
  MXB_CATCH
}
// Copies the pending profile records into buf (text, see Engine::DrainProfile) and clears them.  If cap is too small
// nothing is consumed from the caller's point of view: *needed tells the size and the records are kept for the retry.
MXB_API int MXBEngineDrainProfile(void* h, char* buf, int64_t cap, int64_t* needed) {
  MXB_TRY
  static thread_local std::string stash;
  if (stash.empty()) stash = reinterpret_cast<Engine*>(h)->DrainProfile();
  *needed = (int64_t)stash.size();
  if ((int64_t)stash.size() <= cap) {
    memcpy(buf, stash.data(), stash.size());
    stash.clear();
  }
  MXB_CATCH
}
MXB_API int MXBEngineWaitForVar(void* h, void* var) {
  MXB_TRY
  reinterpret_cast<Engine*>(h)->WaitForVar(reinterpret_cast<Var*>(var));
  MXB_CATCH
}
MXB_API int MXBEngineWaitAll(void* h) {
  MXB_TRY
  reinterpret_cast<Engine*>(h)->WaitAll();
  MXB_CATCH
}
MXB_API int MXBEngineStop(void* h) {
  MXB_TRY
  reinterpret_cast<Engine*>(h)->Stop();
  MXB_CATCH
}
MXB_API int MXBEngineStart(void* h) {
  MXB_TRY
  reinterpret_cast<Engine*>(h)->Start();
  MXB_CATCH
}
MXB_API int MXBEnginePending(void* h, int64_t* out) {
  MXB_TRY
  *out = reinterpret_cast<Engine*>(h)->Pending();
  MXB_CATCH
}
