// Control + data plane for the distributed / elastic kvstore: scheduler, servers and workers over TCP.
//
// Capability parity with the fork's ps-lite (3rdparty/ps-lite): Postoffice (roles, groups, ids
// worker = 9+2r / server = 8+2r / scheduler = 1, barriers, mutable num_workers, dense re-ranking;
// src/postoffice.cc:33-195,329-408), Van (ADD_NODE / BARRIER / MEMBERSHIP_CHANGE_BARRIER / UPDATE_ENV_VAR /
// HEARTBEAT / TERMINATE dispatch, late ADD_NODE of elastic workers, sender dropping; src/van.cc:38-315,555-607),
// ETDefaultNodeManager (host-file diff, removal before addition, launch of new workers, `<hostfile>_log`
// audit lines; src/elastic_training.cc:9-157), KVWorker/KVServer request routing (include/ps/kv_app.h) and
// the Resender's ACK/dedup/retry with PS_DROP_MSG fault injection (src/resender.h, van.cc:430-432,563-570).
//
// Design differences: plain POSIX TCP sockets with length-prefixed frames instead of ZeroMQ + protobuf; one
// receiver thread per node multiplexing all peers with poll(); tensor payloads are sent straight from the
// caller's buffer (scatter-gather writev) and received straight into the destination buffer.  The server-side
// *storage / optimizer* logic lives in python (mxb200/kvstore/dist.py) exactly where the reference runs its
// pickled python optimizer; this file is transport + membership.
#include <arpa/inet.h>
#include <errno.h>
#include <fcntl.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/uio.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../c_api.h"

namespace mxb {
namespace ps {

enum Role : int { kScheduler = 0, kServer = 1, kWorker = 2 };
enum Group : int { kSchedulerGroup = 1, kServerGroup = 2, kWorkerGroup = 4 };  // ps/base.h:20-30
enum Cmd : uint32_t {
  kEmpty = 0, kTerminate, kAddNode, kBarrier, kAck, kHeartbeat, kUpdateEnvVar, kMembershipChangeBarrier,  // message.h:123
  kNodeTable, kBarrierDone, kUpdateEnvAck, kHello, kData, kDataResp, kCommand, kCommandResp,
  kAllGather, kAllGatherDone   // small-blob all-gather among workers through the scheduler (IPC handle exchange)
};

static constexpr uint32_t kMagic = 0x4d584232;  // "MXB2"

struct Header {
  uint32_t magic = kMagic;
  uint32_t cmd = kEmpty;
  int32_t sender = -1;
  int32_t recver = -1;
  int32_t timestamp = -1;
  int32_t push = 0;        // data: 1 push / 0 pull
  int32_t head = 0;        // command head / barrier group / data request code
  int32_t dtype = 0;
  int64_t key = 0;
  int64_t offset = 0;      // element offset of this slice inside the full value
  int64_t total = 0;       // total elements of the full value
  uint64_t sig = 0;        // message signature for ACK / dedup (0 = not tracked)
  uint32_t meta_len = 0;
  uint64_t body_len = 0;
};

struct NodeInfo {
  int id = -1;
  int role = kWorker;
  std::string host;
  int port = 0;
  std::string tag;         // logical host name used by the elastic host file
};

static std::string pack_nodes(const std::vector<NodeInfo>& v) {
  std::ostringstream os;
  for (auto& n : v) os << n.id << ' ' << n.role << ' ' << n.host << ' ' << n.port << ' ' << (n.tag.empty() ? "-" : n.tag) << '\n';
  return os.str();
}
static std::vector<NodeInfo> unpack_nodes(const std::string& s) {
  std::vector<NodeInfo> v;
  std::istringstream is(s);
  NodeInfo n;
  while (is >> n.id >> n.role >> n.host >> n.port >> n.tag) {
    if (n.tag == "-") n.tag.clear();
    v.push_back(n);
  }
  return v;
}

static bool write_all(int fd, const void* p, size_t n) {
  const char* c = static_cast<const char*>(p);
  while (n) {
    ssize_t w = ::send(fd, c, n, MSG_NOSIGNAL);
    if (w < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    c += w;
    n -= (size_t)w;
  }
  return true;
}
static bool read_all(int fd, void* p, size_t n) {
  char* c = static_cast<char*>(p);
  while (n) {
    ssize_t r = ::recv(fd, c, n, 0);
    if (r == 0) return false;
    if (r < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    c += r;
    n -= (size_t)r;
  }
  return true;
}

typedef void (*DataHandler)(void* user, int is_push, int sender, int64_t key, int dtype, int64_t offset, int64_t total,
                            const void* data, uint64_t nbytes, int timestamp, int head);
typedef void (*CommandHandler)(void* user, int head, const char* body, int sender, int timestamp);
typedef int (*LaunchCallback)(void* user, const char* tag, const char* env_csv);

struct PendingPull {
  char* dst = nullptr;
  size_t elem_size = 0;
};

class Node {
 public:
  Node(int role, const std::string& root_host, int root_port, int num_workers, int num_servers, const std::string& tag,
       bool is_new_worker, const std::string& host_file, int base_workers)
      : role_(role), root_host_(root_host), root_port_(root_port), num_workers_(num_workers), num_servers_(num_servers),
        tag_(tag), is_new_worker_(is_new_worker), host_file_(host_file), base_workers_(base_workers) {
    const char* drop = getenv("PS_DROP_MSG");
    drop_pct_ = drop ? atoi(drop) : 0;
    const char* rs = getenv("PS_RESEND");
    resend_ = rs && atoi(rs) != 0;
    const char* rt = getenv("PS_RESEND_TIMEOUT");
    resend_timeout_ms_ = rt ? atoi(rt) : 1000;
    const char* hb = getenv("PS_HEARTBEAT_INTERVAL");
    heartbeat_s_ = hb ? atoi(hb) : 0;
  }
  ~Node() { Stop(); }

  // ------------------------------------------------------------------------------------------ start-up
  void Start() {
    listen_fd_ = ::socket(AF_INET, SOCK_STREAM, 0);
    int one = 1;
    setsockopt(listen_fd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    sockaddr_in addr{};
    addr.sin_family = AF_INET;
    addr.sin_addr.s_addr = htonl(INADDR_ANY);
    addr.sin_port = htons(role_ == kScheduler ? root_port_ : 0);
    if (::bind(listen_fd_, (sockaddr*)&addr, sizeof(addr)) != 0) throw std::runtime_error("ps: bind failed: " + std::string(strerror(errno)));
    ::listen(listen_fd_, 128);
    socklen_t len = sizeof(addr);
    getsockname(listen_fd_, (sockaddr*)&addr, &len);
    my_port_ = ntohs(addr.sin_port);
    running_ = true;
    if (role_ == kScheduler) {
      my_id_ = 1;
      LoadHostFile(&et_workers_);
      recv_thread_ = std::thread([this] { RecvLoop(); });
    } else {
      int fd = ConnectTo(root_host_, root_port_);
      AddConn(1, fd);
      recv_thread_ = std::thread([this] { RecvLoop(); });
      Header h;
      h.cmd = kAddNode;
      h.head = is_new_worker_ ? 1 : 0;
      NodeInfo me;
      me.id = -1; me.role = role_; me.host = "127.0.0.1"; me.port = my_port_; me.tag = tag_;
      const char* nh = getenv("DMLC_NODE_HOST");
      if (nh) me.host = nh;
      Send(1, h, pack_nodes({me}), nullptr, 0);
      // wait for the node table (contains my id)
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [this] { return table_ready_; });
    }
    if (heartbeat_s_ > 0 && role_ != kScheduler) hb_thread_ = std::thread([this] { HeartbeatLoop(); });
    if (resend_) resend_thread_ = std::thread([this] { ResendLoop(); });
  }

  void Stop() {
    if (!running_.exchange(false)) return;
    ::shutdown(listen_fd_, SHUT_RDWR);
    ::close(listen_fd_);
    {
      std::lock_guard<std::mutex> lk(conn_mu_);
      for (auto& kv : conns_) { ::shutdown(kv.second, SHUT_RDWR); ::close(kv.second); }
      for (int fd : anon_) { ::shutdown(fd, SHUT_RDWR); ::close(fd); }
      conns_.clear();
      anon_.clear();
    }
    cv_.notify_all();
    if (recv_thread_.joinable()) recv_thread_.join();
    if (hb_thread_.joinable()) hb_thread_.join();
    if (resend_thread_.joinable()) resend_thread_.join();
  }

  // ------------------------------------------------------------------------------------------ queries
  int my_id() const { return my_id_; }
  int role() const { return role_; }
  int num_workers() {
    std::lock_guard<std::mutex> lk(mu_);
    return num_workers_;
  }
  int num_servers() const { return num_servers_; }
  bool is_new_worker() const { return is_new_worker_; }
  // rank = position of my id in the sorted set of live ids of my role (van.cc:519-539): dense after removals
  int my_rank() {
    std::lock_guard<std::mutex> lk(mu_);
    if (role_ == kScheduler) return 0;
    const std::set<int>& s = role_ == kWorker ? worker_ids_ : server_ids_;
    int r = 0;
    for (int id : s) {
      if (id == my_id_) return r;
      ++r;
    }
    return -1;
  }
  std::vector<int> server_ids() {
    std::lock_guard<std::mutex> lk(mu_);
    return std::vector<int>(server_ids_.begin(), server_ids_.end());
  }
  bool terminated() const { return terminated_; }

  // ------------------------------------------------------------------------------------------ barriers
  void Barrier(int group) {
    int gen;
    {
      std::lock_guard<std::mutex> lk(mu_);
      gen = barrier_gen_;
    }
    Header h;
    h.cmd = kBarrier;
    h.head = group;
    if (role_ == kScheduler) {
      OnBarrier(my_id_, group, false, "");
    } else {
      Send(1, h, "", nullptr, 0);
    }
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return barrier_gen_ != gen || !running_ || terminated_; });
  }

  // Elastic epoch-boundary barrier; kv = "key\nvalue\nkey\nvalue..." forwarded as env to new workers.
  void MembershipChangeBarrier(const std::string& kv) {
    int gen;
    {
      std::lock_guard<std::mutex> lk(mu_);
      gen = barrier_gen_;
    }
    Header h;
    h.cmd = kMembershipChangeBarrier;
    h.head = kWorkerGroup;
    Send(1, h, kv, nullptr, 0);
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return barrier_gen_ != gen || !running_ || terminated_; });
  }

  // All-gather of one small text blob per worker (in rank order) through the scheduler.  Used by the single-box
  // elastic path to exchange CUDA IPC handles of the symmetric arenas after every membership change, when the
  // worker set (and therefore any fixed communicator) has changed.  Blobs must not contain '\n'.
  std::string AllGather(const std::string& blob) {
    int gen;
    {
      std::lock_guard<std::mutex> lk(mu_);
      gen = gather_gen_;
    }
    Header h;
    h.cmd = kAllGather;
    h.head = kWorkerGroup;
    Send(1, h, blob, nullptr, 0);
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return gather_gen_ != gen || !running_ || terminated_; });
    return gather_result_;
  }

  // ------------------------------------------------------------------------------------------ data plane (worker)
  int NewRequest(int expected) {
    std::lock_guard<std::mutex> lk(mu_);
    int ts = next_ts_++;
    tracker_[ts] = {expected, 0};
    return ts;
  }
  void SendData(int recver, int ts, bool push, int64_t key, int dtype, int64_t offset, int64_t total, int head,
                const void* data, uint64_t nbytes, char* pull_dst, size_t elem_size) {
    Header h;
    h.cmd = kData;
    h.timestamp = ts;
    h.push = push ? 1 : 0;
    h.key = key;
    h.dtype = dtype;
    h.offset = offset;
    h.total = total;
    h.head = head;
    if (!push) {
      std::lock_guard<std::mutex> lk(mu_);
      pulls_[((uint64_t)ts << 32) | (uint32_t)recver] = {pull_dst, elem_size};
    }
    Send(recver, h, "", data, nbytes, /*track=*/true);
  }
  void WaitRequest(int ts) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] {
      auto it = tracker_.find(ts);
      return it == tracker_.end() || it->second.second >= it->second.first || !running_ || terminated_;
    });
    tracker_.erase(ts);
  }
  // server -> worker response (pull payload or push ack)
  void Respond(int recver, int ts, int64_t key, const void* data, uint64_t nbytes) {
    Header h;
    h.cmd = kDataResp;
    h.timestamp = ts;
    h.key = key;
    Send(recver, h, "", data, nbytes, /*track=*/true);
  }
  // SimpleApp::Request analogue: (head, body) to every server, wait for all responses
  void SendCommandToServers(int head, const std::string& body) {
    auto servers = server_ids();
    int ts = NewRequest((int)servers.size());
    for (int s : servers) {
      Header h;
      h.cmd = kCommand;
      h.head = head;
      h.timestamp = ts;
      Send(s, h, body, nullptr, 0, true);
    }
    WaitRequest(ts);
  }
  void RespondCommand(int recver, int ts) {
    Header h;
    h.cmd = kCommandResp;
    h.timestamp = ts;
    Send(recver, h, "", nullptr, 0, true);
  }

  void set_data_handler(DataHandler f, void* u) { data_handler_ = f; data_user_ = u; }
  void set_command_handler(CommandHandler f, void* u) { cmd_handler_ = f; cmd_user_ = u; }
  void set_launch_callback(LaunchCallback f, void* u) { launch_cb_ = f; launch_user_ = u; }

  int NumDeadNodes(int timeout_s) {
    std::lock_guard<std::mutex> lk(mu_);
    auto now = std::chrono::steady_clock::now();
    int dead = 0;
    for (auto& kv : last_heartbeat_) {
      if (std::chrono::duration_cast<std::chrono::seconds>(now - kv.second).count() > timeout_s) ++dead;
    }
    return dead;
  }

 private:
  // ------------------------------------------------------------------------------------------ sockets
  int ConnectTo(const std::string& host, int port) {
    for (int attempt = 0; attempt < 600; ++attempt) {
      addrinfo hints{}, *res = nullptr;
      hints.ai_family = AF_INET;
      hints.ai_socktype = SOCK_STREAM;
      if (getaddrinfo(host.c_str(), std::to_string(port).c_str(), &hints, &res) == 0 && res) {
        int fd = ::socket(AF_INET, SOCK_STREAM, 0);
        int one = 1;
        setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
        if (::connect(fd, res->ai_addr, res->ai_addrlen) == 0) {
          freeaddrinfo(res);
          return fd;
        }
        ::close(fd);
        freeaddrinfo(res);
      }
      std::this_thread::sleep_for(std::chrono::milliseconds(100));
      if (!running_) break;
    }
    throw std::runtime_error("ps: cannot connect to " + host + ":" + std::to_string(port));
  }
  void AddConn(int id, int fd) {
    std::lock_guard<std::mutex> lk(conn_mu_);
    conns_[id] = fd;
    send_mu_[fd];
    anon_.erase(fd);
  }
  int FdOf(int id) {
    std::lock_guard<std::mutex> lk(conn_mu_);
    auto it = conns_.find(id);
    return it == conns_.end() ? -1 : it->second;
  }
  void EnsureConn(const NodeInfo& n) {
    if (n.id == my_id_ || FdOf(n.id) >= 0) return;
    int fd = ConnectTo(n.host, n.port);
    AddConn(n.id, fd);
    Header h;
    h.cmd = kHello;
    Send(n.id, h, "", nullptr, 0);
  }

  void Send(int recver, Header h, const std::string& meta, const void* body, uint64_t nbytes, bool track = false) {
    h.sender = my_id_;
    h.recver = recver;
    h.meta_len = (uint32_t)meta.size();
    h.body_len = nbytes;
    if (track && resend_) {
      h.sig = ((uint64_t)(uint32_t)my_id_ << 40) ^ (sig_counter_.fetch_add(1) + 1);
      std::lock_guard<std::mutex> lk(resend_mu_);
      Unacked u;
      u.h = h;
      u.meta = meta;
      u.body.assign((const char*)body, (const char*)body + nbytes);
      u.sent = std::chrono::steady_clock::now();
      unacked_[h.sig] = std::move(u);
    }
    RawSend(recver, h, meta, body, nbytes);
  }
  void RawSend(int recver, const Header& h, const std::string& meta, const void* body, uint64_t nbytes) {
    int fd = FdOf(recver);
    if (fd < 0) return;   // dropped / unknown peer (removed worker): silently discard, like a dropped sender
    std::mutex* m;
    {
      std::lock_guard<std::mutex> lk(conn_mu_);
      m = &send_mu_[fd];
    }
    std::lock_guard<std::mutex> lk(*m);
    if (!write_all(fd, &h, sizeof(h))) return;
    if (!meta.empty() && !write_all(fd, meta.data(), meta.size())) return;
    if (nbytes) write_all(fd, body, nbytes);
  }

  // ------------------------------------------------------------------------------------------ receive loop
  void RecvLoop() {
    std::mt19937 rng(12345u + (unsigned)my_port_);
    while (running_) {
      std::vector<pollfd> fds;
      std::vector<int> ids;
      fds.push_back({listen_fd_, POLLIN, 0});
      ids.push_back(0);
      {
        std::lock_guard<std::mutex> lk(conn_mu_);
        for (auto& kv : conns_) { fds.push_back({kv.second, POLLIN, 0}); ids.push_back(kv.first); }
        for (int fd : anon_) { fds.push_back({fd, POLLIN, 0}); ids.push_back(-1); }
      }
      int rc = ::poll(fds.data(), fds.size(), 200);
      if (rc <= 0) continue;
      if (fds[0].revents & POLLIN) {
        int fd = ::accept(listen_fd_, nullptr, nullptr);
        if (fd >= 0) {
          int one = 1;
          setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
          std::lock_guard<std::mutex> lk(conn_mu_);
          anon_.insert(fd);
          send_mu_[fd];
        }
      }
      for (size_t i = 1; i < fds.size(); ++i) {
        if (!(fds[i].revents & (POLLIN | POLLHUP | POLLERR))) continue;
        Header h;
        if (!read_all(fds[i].fd, &h, sizeof(h)) || h.magic != kMagic) {
          DropFd(fds[i].fd);
          continue;
        }
        std::string meta(h.meta_len, '\0');
        if (h.meta_len && !read_all(fds[i].fd, &meta[0], h.meta_len)) { DropFd(fds[i].fd); continue; }
        // payload: pulls land directly in the destination buffer
        std::vector<char> body;
        char* dst = nullptr;
        if (h.cmd == kDataResp && h.body_len) {
          std::lock_guard<std::mutex> lk(mu_);
          auto it = pulls_.find(((uint64_t)h.timestamp << 32) | (uint32_t)h.sender);
          if (it != pulls_.end()) dst = it->second.dst;
        }
        if (h.body_len) {
          if (dst) {
            if (!read_all(fds[i].fd, dst, h.body_len)) { DropFd(fds[i].fd); continue; }
          } else {
            body.resize(h.body_len);
            if (!read_all(fds[i].fd, body.data(), h.body_len)) { DropFd(fds[i].fd); continue; }
          }
        }
        if (ids[i] == -1 && h.sender > 0) AddConn(h.sender, fds[i].fd);
        // removed workers are ignored from now on (DropSender / IsSenderIdValid, zmq_van.h:128-156)
        {
          std::lock_guard<std::mutex> lk(mu_);
          if (dropped_ids_.count(h.sender)) continue;
        }
        // fault injection + reliable delivery
        if (h.sig != 0) {
          if (drop_pct_ > 0 && (int)(rng() % 100) < drop_pct_ && h.cmd != kAck) continue;   // PS_DROP_MSG
          Header ack;
          ack.cmd = kAck;
          ack.sig = h.sig;
          Send(h.sender, ack, "", nullptr, 0);
          std::lock_guard<std::mutex> lk(resend_mu_);
          if (!seen_sigs_.insert(h.sig).second) continue;   // duplicate of a resent message
        }
        Dispatch(h, meta, body, fds[i].fd);
      }
    }
  }
  void DropFd(int fd) {
    std::lock_guard<std::mutex> lk(conn_mu_);
    for (auto it = conns_.begin(); it != conns_.end();) {
      if (it->second == fd) it = conns_.erase(it);
      else ++it;
    }
    anon_.erase(fd);
    ::close(fd);
  }

  void Dispatch(const Header& h, const std::string& meta, std::vector<char>& body, int fd) {
    switch (h.cmd) {
      case kAck: {
        std::lock_guard<std::mutex> lk(resend_mu_);
        unacked_.erase(h.sig);
        break;
      }
      case kHello:
        break;
      case kAddNode:
        if (role_ == kScheduler) OnAddNode(h, meta, fd);
        break;
      case kNodeTable:
        OnNodeTable(meta, h.head);
        break;
      case kBarrier:
        if (role_ == kScheduler) OnBarrier(h.sender, h.head, false, "");
        break;
      case kMembershipChangeBarrier:
        if (role_ == kScheduler) OnBarrier(h.sender, h.head, true, meta);
        break;
      case kBarrierDone: {
        std::lock_guard<std::mutex> lk(mu_);
        ++barrier_gen_;
        cv_.notify_all();
        break;
      }
      case kAllGather:
        if (role_ == kScheduler) OnAllGather(h.sender, meta);
        break;
      case kAllGatherDone: {
        std::lock_guard<std::mutex> lk(mu_);
        gather_result_ = meta;
        ++gather_gen_;
        cv_.notify_all();
        break;
      }
      case kUpdateEnvVar:
        OnUpdateEnvVar(h, meta);
        break;
      case kUpdateEnvAck:
        if (role_ == kScheduler) OnUpdateEnvAck();
        break;
      case kHeartbeat:
        if (role_ == kScheduler) {
          std::lock_guard<std::mutex> lk(mu_);
          last_heartbeat_[h.sender] = std::chrono::steady_clock::now();
        }
        break;
      case kTerminate: {
        std::lock_guard<std::mutex> lk(mu_);
        terminated_ = true;
        cv_.notify_all();
        break;
      }
      case kData:
        if (data_handler_)
          data_handler_(data_user_, h.push, h.sender, h.key, h.dtype, h.offset, h.total, body.data(), h.body_len,
                        h.timestamp, h.head);
        break;
      case kDataResp:
      case kCommandResp: {
        std::lock_guard<std::mutex> lk(mu_);
        pulls_.erase(((uint64_t)h.timestamp << 32) | (uint32_t)h.sender);
        auto it = tracker_.find(h.timestamp);
        if (it != tracker_.end()) it->second.second++;
        cv_.notify_all();
        break;
      }
      case kCommand:
        if (cmd_handler_) cmd_handler_(cmd_user_, h.head, meta.c_str(), h.sender, h.timestamp);
        else RespondCommand(h.sender, h.timestamp);
        break;
      default:
        break;
    }
  }

  // ------------------------------------------------------------------------------------------ scheduler: membership
  void OnAddNode(const Header& h, const std::string& meta, int fd) {
    NodeInfo n = unpack_nodes(meta)[0];
    std::vector<NodeInfo> table;
    bool late = false;
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (n.role == kWorker) n.id = 9 + 2 * total_workers_seen_++;
      else n.id = 8 + 2 * total_servers_seen_++;
      nodes_[n.id] = n;
      (n.role == kWorker ? worker_ids_ : server_ids_).insert(n.id);
      last_heartbeat_[n.id] = std::chrono::steady_clock::now();
      late = initial_table_sent_;
      if (n.role == kWorker && !n.tag.empty()) et_workers_.insert(n.tag);
    }
    AddConn(n.id, fd);
    if (late) {
      // elastic newcomer (van.cc:95-185): it gets the whole table, servers learn about it, nobody else needs to connect
      {
        std::lock_guard<std::mutex> lk(mu_);
        for (auto& kv : nodes_) table.push_back(kv.second);
      }
      Header t;
      t.cmd = kNodeTable;
      t.head = n.id;                       // "your id"
      Send(n.id, t, pack_nodes(table) + "#" + std::to_string(num_workers()), nullptr, 0);
      Header t2;
      t2.cmd = kNodeTable;
      t2.head = 0;
      for (int s : server_ids()) Send(s, t2, pack_nodes({n}) + "#" + std::to_string(num_workers()), nullptr, 0);
      return;
    }
    bool complete;
    {
      std::lock_guard<std::mutex> lk(mu_);
      complete = (int)worker_ids_.size() >= num_workers_ && (int)server_ids_.size() >= num_servers_;
      if (complete) {
        initial_table_sent_ = true;
        for (auto& kv : nodes_) table.push_back(kv.second);
      }
    }
    if (complete) {
      for (auto& e : table) {
        Header t;
        t.cmd = kNodeTable;
        t.head = e.id;
        Send(e.id, t, pack_nodes(table) + "#" + std::to_string(num_workers()), nullptr, 0);
      }
    }
  }

  void OnNodeTable(const std::string& meta, int my_assigned_id) {
    // payload: node lines + "#<num_workers>"
    auto pos = meta.rfind('#');
    std::string lines = meta.substr(0, pos);
    int nw = atoi(meta.substr(pos + 1).c_str());
    auto v = unpack_nodes(lines);
    std::vector<NodeInfo> to_connect;
    {
      std::lock_guard<std::mutex> lk(mu_);
      for (auto& n : v) {
        nodes_[n.id] = n;
        (n.role == kWorker ? worker_ids_ : server_ids_).insert(n.id);
        if (my_assigned_id > 0 && n.id == my_assigned_id && my_id_ < 0) my_id_ = n.id;
        // workers <-> servers only (zmq_van.h:103-106)
        if ((role_ == kWorker && n.role == kServer)) to_connect.push_back(n);
      }
      if (nw > 0) num_workers_ = nw;
    }
    for (auto& n : to_connect) EnsureConn(n);
    {
      std::lock_guard<std::mutex> lk(mu_);
      table_ready_ = true;
      cv_.notify_all();
    }
  }

  int GroupSize(int group) {
    int n = 0;
    if (group & kSchedulerGroup) n += 1;
    // expected (not merely registered) sizes: a barrier reached before every node has joined must wait for them
    if (group & kServerGroup) n += std::max((int)server_ids_.size(), num_servers_);
    if (group & kWorkerGroup) n += std::max((int)worker_ids_.size(), num_workers_);
    return n;
  }
  std::vector<int> GroupMembers(int group) {
    std::vector<int> v;
    if (group & kSchedulerGroup) v.push_back(1);
    if (group & kServerGroup) v.insert(v.end(), server_ids_.begin(), server_ids_.end());
    if (group & kWorkerGroup) v.insert(v.end(), worker_ids_.begin(), worker_ids_.end());
    return v;
  }

  void OnAllGather(int sender, const std::string& blob) {
    std::vector<int> members;
    std::string all;
    {
      std::lock_guard<std::mutex> lk(mu_);
      gather_blobs_[sender] = blob;
      if ((int)gather_blobs_.size() < GroupSize(kWorkerGroup)) return;
      for (auto& kv : gather_blobs_) {   // std::map: ascending node id == ascending rank (id = 9 + 2*rank)
        members.push_back(kv.first);
        all += kv.second;
        all += '\n';
      }
      gather_blobs_.clear();
    }
    for (int id : members) {
      Header h;
      h.cmd = kAllGatherDone;
      Send(id, h, all, nullptr, 0);
    }
  }

  void OnBarrier(int sender, int group, bool membership, const std::string& kv) {
    std::vector<int> release;
    bool fire_membership = false;
    int max_id = 0;
    {
      std::lock_guard<std::mutex> lk(mu_);
      int& cnt = barrier_count_[group];
      ++cnt;
      if (cnt < GroupSize(group)) return;
      cnt = 0;
      release = GroupMembers(group);
      if (membership) {
        fire_membership = true;
        for (int id : release) max_id = std::max(max_id, id);
        pending_release_ = release;
        pending_max_id_ = max_id;
        pending_kv_ = kv;
      }
    }
    if (fire_membership) {
      InvokeMembershipChange();
      return;
    }
    ReleaseBarrier(release, 1 << 30);
  }

  void ReleaseBarrier(const std::vector<int>& ids, int max_id) {
    std::set<int> dropped;
    {
      std::lock_guard<std::mutex> lk(mu_);
      dropped = dropped_ids_;
    }
    bool self = false;
    for (int id : ids) {
      if (id > max_id || dropped.count(id)) continue;   // newcomers never see a response they did not ask for
      if (id == my_id_) {
        self = true;
        continue;
      }
      Header h;
      h.cmd = kBarrierDone;
      Send(id, h, "", nullptr, 0);
    }
    // Release this node LAST: its main thread may be sitting in the final barrier before Stop(); waking it first let
    // it tear the sockets down while the loop above was still notifying the other nodes (rare shutdown hang).
    if (self) {
      std::lock_guard<std::mutex> lk(mu_);
      ++barrier_gen_;
      cv_.notify_all();
    }
  }

  // ---- ETDefaultNodeManager (elastic_training.cc) ----
  void LoadHostFile(std::set<std::string>* out) {
    out->clear();
    if (host_file_.empty()) return;
    std::ifstream f(host_file_);
    std::string line;
    while (std::getline(f, line)) {
      while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
      if (!line.empty()) out->insert(line);
    }
  }
  void AppendLog(const char* what, const std::string& tag) {
    if (host_file_.empty()) return;
    std::ofstream f(host_file_ + "_log", std::ios::app);
    auto ns = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
    f << log_seq_++ << ' ' << what << ' ' << tag << ' ' << ns << '\n';
  }

  void InvokeMembershipChange() {
    std::set<std::string> file_set;
    LoadHostFile(&file_set);
    std::vector<std::string> removed, added;
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (!host_file_.empty()) {
        for (auto& w : et_workers_) if (!file_set.count(w)) removed.push_back(w);
        for (auto& w : file_set) if (!et_workers_.count(w)) added.push_back(w);
      }
    }
    // removal takes priority over addition (elastic_training.cc:91-107): a swap is spread over two epochs
    if (!removed.empty()) {
      std::vector<int> victim_ids;
      {
        std::lock_guard<std::mutex> lk(mu_);
        for (auto& kv : nodes_)
          if (kv.second.role == kWorker && worker_ids_.count(kv.first) &&
              std::find(removed.begin(), removed.end(), kv.second.tag) != removed.end())
            victim_ids.push_back(kv.first);
      }
      pending_added_.clear();
      pending_removed_tags_ = removed;
      BroadcastNumWorkers(num_workers() - (int)victim_ids.size(), victim_ids, {});
    } else if (!added.empty()) {
      pending_added_ = added;
      pending_removed_tags_.clear();
      std::vector<int> new_ids;
      {
        std::lock_guard<std::mutex> lk(mu_);
        for (size_t i = 0; i < added.size(); ++i) new_ids.push_back(9 + 2 * (total_workers_seen_ + (int)i));
      }
      BroadcastNumWorkers(num_workers() + (int)added.size(), {}, new_ids);
    } else {
      pending_added_.clear();
      pending_removed_tags_.clear();
      OnSuccessUpdatingEnv();
    }
  }

  // UPDATE_ENV_VAR{DMLC_NUM_WORKER, value, removed ids} to every live node except the victims (postoffice.cc:132-187)
  void BroadcastNumWorkers(int new_count, const std::vector<int>& victims, const std::vector<int>& new_ids) {
    std::vector<int> targets;
    {
      std::lock_guard<std::mutex> lk(mu_);
      for (int id : server_ids_) targets.push_back(id);
      for (int id : worker_ids_)
        if (std::find(victims.begin(), victims.end(), id) == victims.end()) targets.push_back(id);
      update_acks_pending_ = (int)targets.size();
    }
    std::ostringstream os;
    os << "DMLC_NUM_WORKER\n" << new_count << "\n";
    for (size_t i = 0; i < victims.size(); ++i) os << (i ? "," : "") << victims[i];
    os << "\n";
    for (size_t i = 0; i < new_ids.size(); ++i) os << (i ? "," : "") << new_ids[i];
    ApplyNumWorkers(new_count, victims, new_ids);
    if (targets.empty()) {
      OnSuccessUpdatingEnv();
      return;
    }
    for (int id : targets) {
      Header h;
      h.cmd = kUpdateEnvVar;
      Send(id, h, os.str(), nullptr, 0);
    }
    // tell the victims to go away (the reference leaves them blocked forever: "TODO send terminate command")
    for (int id : victims) {
      Header h;
      h.cmd = kTerminate;
      Send(id, h, "", nullptr, 0);
    }
  }
  // updateNumWorker (postoffice.cc:71-119): drop the victims, pre-register the ids of the expected newcomers so that
  // group sizes (barriers, server-side sync counts) include them from this moment on
  void ApplyNumWorkers(int new_count, const std::vector<int>& victims, const std::vector<int>& new_ids) {
    std::lock_guard<std::mutex> lk(mu_);
    num_workers_ = new_count;
    for (int id : new_ids) worker_ids_.insert(id);
    for (int id : victims) {
      worker_ids_.erase(id);
      dropped_ids_.insert(id);
      auto it = nodes_.find(id);
      if (it != nodes_.end()) et_workers_.erase(it->second.tag);
    }
  }
  void OnUpdateEnvVar(const Header& h, const std::string& meta) {
    std::istringstream is(meta);
    std::string name, val, ids, adds;
    std::getline(is, name);
    std::getline(is, val);
    std::getline(is, ids);
    std::getline(is, adds);
    if (name == "DMLC_NUM_WORKER") {
      std::vector<int> victims;
      std::istringstream vs(ids);
      std::string tok;
      while (std::getline(vs, tok, ',')) if (!tok.empty()) victims.push_back(atoi(tok.c_str()));
      std::vector<int> new_ids;
      std::istringstream as(adds);
      while (std::getline(as, tok, ',')) if (!tok.empty()) new_ids.push_back(atoi(tok.c_str()));
      ApplyNumWorkers(atoi(val.c_str()), victims, new_ids);
      setenv("DMLC_NUM_WORKER", val.c_str(), 1);
    }
    Header a;
    a.cmd = kUpdateEnvAck;
    Send(1, a, "", nullptr, 0);
  }
  void OnUpdateEnvAck() {
    bool done;
    {
      std::lock_guard<std::mutex> lk(mu_);
      done = --update_acks_pending_ == 0;
    }
    if (done) OnSuccessUpdatingEnv();
  }
  void OnSuccessUpdatingEnv() {
    for (auto& t : pending_removed_tags_) AppendLog("REMOVED", t);
    for (auto& t : pending_added_) {
      LaunchNewWorker(t);
      AppendLog("ADDED", t);
    }
    std::vector<int> rel;
    int max_id;
    {
      std::lock_guard<std::mutex> lk(mu_);
      rel = pending_release_;
      max_id = pending_max_id_;
    }
    ReleaseBarrier(rel, max_id);
  }
  // launchCommandOnNewWorker (elastic_training.cc:26-62): env for the newcomer = NEW_WORKER, counts, root uri/port and
  // the barrier payload (EPOCH_BEGIN) verbatim.
  void LaunchNewWorker(const std::string& tag) {
    std::ostringstream env;
    env << "NEW_WORKER=1,DMLC_ROLE=worker,DMLC_NUM_WORKER=" << num_workers() << ",DMLC_NUM_SERVER=" << num_servers_
        << ",DMLC_PS_ROOT_URI=" << root_host_ << ",DMLC_PS_ROOT_PORT=" << root_port_ << ",ELASTIC_TRAINING_ENABLED=1"
        << ",MXB200_WORKER_TAG=" << tag;
    std::istringstream is(pending_kv_);
    std::string k, v;
    while (std::getline(is, k) && std::getline(is, v)) env << "," << k << "=" << v;
    if (launch_cb_) {
      launch_cb_(launch_user_, tag.c_str(), env.str().c_str());
      return;
    }
    const char* script = getenv("MXNET_LAUNCH_SCRIPT_PATH");
    const char* cmd = getenv("TRAINING_CMD");
    if (!script || !cmd) return;
    std::ostringstream c;
    c << "python " << script << " -n 1 --host " << tag << " --launch-worker True";
    std::istringstream es(env.str());
    std::string kvp;
    while (std::getline(es, kvp, ',')) {
      auto eq = kvp.find('=');
      c << " --env " << kvp.substr(0, eq) << ":" << kvp.substr(eq + 1);
    }
    c << " " << cmd << " &";
    int rc = system(c.str().c_str());
    (void)rc;
  }

  // ------------------------------------------------------------------------------------------ background threads
  void HeartbeatLoop() {
    while (running_) {
      for (int i = 0; i < heartbeat_s_ * 10 && running_; ++i) std::this_thread::sleep_for(std::chrono::milliseconds(100));
      Header h;
      h.cmd = kHeartbeat;
      Send(1, h, "", nullptr, 0);
    }
  }
  struct Unacked {
    Header h;
    std::string meta;
    std::vector<char> body;
    std::chrono::steady_clock::time_point sent;
    int retries = 0;
  };
  void ResendLoop() {
    while (running_) {
      std::this_thread::sleep_for(std::chrono::milliseconds(std::max(10, resend_timeout_ms_ / 4)));
      std::vector<Unacked> again;
      {
        std::lock_guard<std::mutex> lk(resend_mu_);
        auto now = std::chrono::steady_clock::now();
        for (auto& kv : unacked_) {
          auto& u = kv.second;
          auto age = std::chrono::duration_cast<std::chrono::milliseconds>(now - u.sent).count();
          if (age > (long)resend_timeout_ms_ * (1 + u.retries) && u.retries < 10) {
            u.retries++;
            u.sent = now;
            again.push_back(u);
          }
        }
      }
      for (auto& u : again) RawSend(u.h.recver, u.h, u.meta, u.body.data(), u.body.size());
    }
  }

  int role_;
  std::string root_host_;
  int root_port_;
  int num_workers_, num_servers_;
  std::string tag_;
  bool is_new_worker_;
  std::string host_file_;
  int base_workers_;
  int my_id_ = -1, my_port_ = 0, listen_fd_ = -1;
  std::atomic<bool> running_{false};
  bool terminated_ = false;
  std::thread recv_thread_, hb_thread_, resend_thread_;
  std::mutex mu_, conn_mu_, resend_mu_;
  std::condition_variable cv_;
  std::map<int, int> conns_;
  std::set<int> anon_;
  std::map<int, std::mutex> send_mu_;
  std::map<int, NodeInfo> nodes_;
  std::set<int> worker_ids_, server_ids_, dropped_ids_;
  int total_workers_seen_ = 0, total_servers_seen_ = 0;
  bool table_ready_ = false, initial_table_sent_ = false;
  std::map<int, int> barrier_count_;
  int barrier_gen_ = 0;
  std::vector<int> pending_release_;
  int pending_max_id_ = 0;
  std::string pending_kv_;
  std::map<int, std::string> gather_blobs_;
  std::string gather_result_;
  int gather_gen_ = 0;
  std::vector<std::string> pending_added_, pending_removed_tags_;
  std::set<std::string> et_workers_;
  int update_acks_pending_ = 0;
  int log_seq_ = 0;
  int next_ts_ = 0;
  std::map<int, std::pair<int, int>> tracker_;
  std::map<uint64_t, PendingPull> pulls_;
  std::map<int, std::chrono::steady_clock::time_point> last_heartbeat_;
  DataHandler data_handler_ = nullptr;
  void* data_user_ = nullptr;
  CommandHandler cmd_handler_ = nullptr;
  void* cmd_user_ = nullptr;
  LaunchCallback launch_cb_ = nullptr;
  void* launch_user_ = nullptr;
  int drop_pct_ = 0;
  bool resend_ = false;
  int resend_timeout_ms_ = 1000;
  int heartbeat_s_ = 0;
  std::atomic<uint64_t> sig_counter_{0};
  std::map<uint64_t, Unacked> unacked_;
  std::unordered_set<uint64_t> seen_sigs_;
};

}  // namespace ps
}  // namespace mxb

using mxb::ps::Node;

MXB_API int MXBPsCreate(int role, const char* root_host, int root_port, int num_workers, int num_servers,
                        const char* tag, int is_new_worker, const char* host_file, int base_workers, void** out) {
  MXB_TRY
  *out = new Node(role, root_host ? root_host : "127.0.0.1", root_port, num_workers, num_servers, tag ? tag : "",
                  is_new_worker != 0, host_file ? host_file : "", base_workers);
  MXB_CATCH
}
MXB_API int MXBPsStart(void* h) {
  MXB_TRY
  reinterpret_cast<Node*>(h)->Start();
  MXB_CATCH
}
MXB_API int MXBPsStop(void* h) {
  MXB_TRY
  reinterpret_cast<Node*>(h)->Stop();
  MXB_CATCH
}
MXB_API int MXBPsFree(void* h) {
  MXB_TRY
  delete reinterpret_cast<Node*>(h);
  MXB_CATCH
}
MXB_API int MXBPsInfo(void* h, int* my_id, int* rank, int* num_workers, int* num_servers, int* terminated) {
  MXB_TRY
  Node* n = reinterpret_cast<Node*>(h);
  *my_id = n->my_id();
  *rank = n->my_rank();
  *num_workers = n->num_workers();
  *num_servers = n->num_servers();
  *terminated = n->terminated() ? 1 : 0;
  MXB_CATCH
}
MXB_API int MXBPsServerIds(void* h, int* out, int cap, int* n) {
  MXB_TRY
  auto v = reinterpret_cast<Node*>(h)->server_ids();
  *n = (int)v.size();
  for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
  MXB_CATCH
}
MXB_API int MXBPsBarrier(void* h, int group) {
  MXB_TRY
  reinterpret_cast<Node*>(h)->Barrier(group);
  MXB_CATCH
}
MXB_API int MXBPsMembershipChangeBarrier(void* h, const char* kv) {
  MXB_TRY
  reinterpret_cast<Node*>(h)->MembershipChangeBarrier(kv ? kv : "");
  MXB_CATCH
}
// out receives the '\n'-separated blobs of all workers in rank order (truncated to cap-1 bytes + NUL); *out_len = full length
MXB_API int MXBPsAllGather(void* h, const char* blob, char* out, int cap, int* out_len) {
  MXB_TRY
  std::string r = reinterpret_cast<Node*>(h)->AllGather(blob ? blob : "");
  *out_len = (int)r.size();
  if (cap > 0) {
    const int n = std::min((int)r.size(), cap - 1);
    memcpy(out, r.data(), n);
    out[n] = 0;
  }
  MXB_CATCH
}
MXB_API int MXBPsNewRequest(void* h, int expected, int* ts) {
  MXB_TRY
  *ts = reinterpret_cast<Node*>(h)->NewRequest(expected);
  MXB_CATCH
}
MXB_API int MXBPsSendData(void* h, int recver, int ts, int push, int64_t key, int dtype, int64_t offset, int64_t total,
                          int head, const void* data, uint64_t nbytes, void* pull_dst, uint64_t elem_size) {
  MXB_TRY
  reinterpret_cast<Node*>(h)->SendData(recver, ts, push != 0, key, dtype, offset, total, head, data, nbytes,
                                       reinterpret_cast<char*>(pull_dst), (size_t)elem_size);
  MXB_CATCH
}
MXB_API int MXBPsWait(void* h, int ts) {
  MXB_TRY
  reinterpret_cast<Node*>(h)->WaitRequest(ts);
  MXB_CATCH
}
MXB_API int MXBPsRespond(void* h, int recver, int ts, int64_t key, const void* data, uint64_t nbytes) {
  MXB_TRY
  reinterpret_cast<Node*>(h)->Respond(recver, ts, key, data, nbytes);
  MXB_CATCH
}
MXB_API int MXBPsSendCommand(void* h, int head, const char* body) {
  MXB_TRY
  reinterpret_cast<Node*>(h)->SendCommandToServers(head, body ? body : "");
  MXB_CATCH
}
MXB_API int MXBPsRespondCommand(void* h, int recver, int ts) {
  MXB_TRY
  reinterpret_cast<Node*>(h)->RespondCommand(recver, ts);
  MXB_CATCH
}
MXB_API int MXBPsSetDataHandler(void* h, mxb::ps::DataHandler f, void* user) {
  MXB_TRY
  reinterpret_cast<Node*>(h)->set_data_handler(f, user);
  MXB_CATCH
}
MXB_API int MXBPsSetCommandHandler(void* h, mxb::ps::CommandHandler f, void* user) {
  MXB_TRY
  reinterpret_cast<Node*>(h)->set_command_handler(f, user);
  MXB_CATCH
}
MXB_API int MXBPsSetLaunchCallback(void* h, mxb::ps::LaunchCallback f, void* user) {
  MXB_TRY
  reinterpret_cast<Node*>(h)->set_launch_callback(f, user);
  MXB_CATCH
}
MXB_API int MXBPsNumDeadNodes(void* h, int timeout_s, int* out) {
  MXB_TRY
  *out = reinterpret_cast<Node*>(h)->NumDeadNodes(timeout_s);
  MXB_CATCH
}
