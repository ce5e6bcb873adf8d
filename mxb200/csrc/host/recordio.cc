// Native RecordIO reader + prefetching batch loader (the first half of the reference's C++ image pipeline:
// dmlc::RecordIOReader / InputSplit, src/io/iter_image_recordio_2.cc:79-236 (chunk parse, part_index / num_parts
// sharding, shuffle), src/io/iter_batchloader.h (batch assembly, round_batch padding), src/io/iter_prefetcher.h
// (bounded queue of ready batches)).
//
// Design for this machine: the expensive half of the pipeline (JPEG decode, resize, crop, mirror, normalise) runs on
// the GPU (csrc/cuda/jpeg.cu: nvJPEG batched decode + one fused augmentation kernel), so the host side never touches
// pixels.  What the host must do fast is (1) find records, (2) strip framing and image headers, (3) pack the COMPRESSED
// payloads of a batch back to back into a pinned staging slot handed in by the caller, (4) keep `depth` such slots
// filled ahead of the consumer.  The file is mmap'ed once; worker threads memcpy out of the page cache.
//
//   record  := u32 magic 0xced7230a | u32 (cflag << 29 | length) | payload | pad to 4 bytes
//   cflag   := 0 whole record, 1 first part, 2 middle part, 3 last part (a payload containing the magic is split there
//              and the magic word is re-inserted between the parts on read)
//   payload := IRHeader { u32 flag; f32 label; u64 id; u64 id2 } | flag * f32 extra labels (flag > 0) | image bytes
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <fstream>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "../c_api.h"

namespace {

constexpr uint32_t kMagic = 0xced7230a;

struct Rec {
  int64_t off;        // file offset of the first part's magic word
  int64_t len;        // logical payload length (parts joined, magic words re-inserted)
  int32_t parts;      // 1 = contiguous
};

struct RecFile {
  int fd = -1;
  const uint8_t* base = nullptr;
  int64_t size = 0;
  std::vector<Rec> recs;
  std::vector<int64_t> keys;     // keys of the .idx file (or 0..n-1)

  ~RecFile() {
    if (base) munmap(const_cast<uint8_t*>(base), (size_t)size);
    if (fd >= 0) close(fd);
  }

  static uint32_t rd32(const uint8_t* p) {
    uint32_t v;
    memcpy(&v, p, 4);
    return v;
  }

  // parse the record that starts at `off`; returns the offset just past it
  int64_t parse(int64_t off, Rec* r) const {
    r->off = off;
    r->len = 0;
    r->parts = 0;
    for (;;) {
      MXB_CHECK(off + 8 <= size, "RecordIO: truncated record header");
      MXB_CHECK(rd32(base + off) == kMagic, "RecordIO: bad magic (corrupt file or wrong offset)");
      const uint32_t lrec = rd32(base + off + 4);
      const uint32_t cflag = lrec >> 29, len = lrec & ((1u << 29) - 1);
      MXB_CHECK(off + 8 + (int64_t)len <= size, "RecordIO: truncated record payload");
      r->len += len + (r->parts > 0 ? 4 : 0);
      r->parts += 1;
      off += 8 + (((int64_t)len + 3) & ~(int64_t)3);
      if (cflag == 0 || cflag == 3) return off;
    }
  }

  void open(const char* path, const char* idx_path) {
    fd = ::open(path, O_RDONLY);
    MXB_CHECK(fd >= 0, std::string("cannot open ") + path);
    struct stat st;
    MXB_CHECK(fstat(fd, &st) == 0, "fstat failed");
    size = st.st_size;
    if (size > 0) {
      void* m = mmap(nullptr, (size_t)size, PROT_READ, MAP_SHARED, fd, 0);
      MXB_CHECK(m != MAP_FAILED, "mmap failed");
      base = static_cast<const uint8_t*>(m);
      madvise(m, (size_t)size, MADV_WILLNEED);
    }
    if (idx_path && idx_path[0]) {
      std::ifstream f(idx_path);
      MXB_CHECK(f.good(), std::string("cannot open ") + idx_path);
      int64_t key, off;
      while (f >> key >> off) {
        Rec r;
        parse(off, &r);
        recs.push_back(r);
        keys.push_back(key);
      }
    } else {
      int64_t off = 0;
      while (off < size) {
        Rec r;
        off = parse(off, &r);
        keys.push_back((int64_t)recs.size());
        recs.push_back(r);
      }
    }
  }

  // copy the logical payload bytes [skip, len) of record i to dst
  void copy(const Rec& r, int64_t skip, uint8_t* dst) const {
    int64_t off = r.off, logical = 0;
    for (int part = 0; part < r.parts; ++part) {
      const uint32_t len = rd32(base + off + 4) & ((1u << 29) - 1);
      if (part > 0) {     // the magic word that split the payload
        const uint8_t m[4] = {0x0a, 0x23, 0xd7, 0xce};
        for (int b = 0; b < 4; ++b, ++logical)
          if (logical >= skip) dst[logical - skip] = m[b];
      }
      const int64_t from = std::max<int64_t>(0, skip - logical);
      if (from < (int64_t)len) memcpy(dst + (logical + from - skip), base + off + 8 + from, (size_t)(len - from));
      logical += len;
      off += 8 + (((int64_t)len + 3) & ~(int64_t)3);
    }
  }
};

struct ImageHeader {
  uint32_t flag;
  float label;
  uint64_t id, id2;
};
static_assert(sizeof(ImageHeader) == 24, "IRHeader layout");

struct Slot {
  uint8_t* blob = nullptr;
  std::vector<int64_t> offsets;    // count + 1 (16-byte aligned starts)
  std::vector<int64_t> lengths;    // exact payload bytes
  std::vector<float> labels;       // count * label_width
  std::vector<int64_t> ids;
  int count = 0, pad = 0;
  int64_t batch_index = -1;        // which batch of the epoch lives here
  bool ready = false;
  std::string error;
};

struct Loader {
  RecFile* rec;
  int batch, label_width, depth;
  int64_t slot_bytes;
  std::vector<Slot> slots;
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_ready, cv_free;
  // epoch state (guarded by mu)
  std::vector<int64_t> order;      // record indices of this rank's shard in visiting order
  int64_t num_batches = 0, next_claim = 0, next_deliver = 0, released_upto = 0;
  uint64_t epoch = 0;
  bool round_batch = true, stop = false;

  void fill(Slot& s, int64_t b, const std::vector<int64_t>& ord, bool round) {
    const int64_t n = (int64_t)ord.size();
    const int64_t first = b * batch;
    s.count = (int)std::min<int64_t>(batch, round ? batch : n - first);
    s.pad = (int)std::max<int64_t>(0, first + s.count - n);
    s.offsets.assign(s.count + 1, 0);
    s.labels.assign((size_t)s.count * label_width, 0.f);
    s.ids.assign(s.count, 0);
    s.lengths.assign(s.count, 0);
    int64_t pos = 0;
    for (int i = 0; i < s.count; ++i) {
      const Rec& r = rec->recs[ord[(first + i) % n]];     // round_batch: wrap to the start of the shard
      MXB_CHECK(r.len >= (int64_t)sizeof(ImageHeader), "RecordIO: record shorter than an image header");
      // a split record (its payload contained the magic word) is joined once; the common case reads the mapping in place
      std::vector<uint8_t> joined;
      const uint8_t* payload = rec->base + r.off + 8;
      if (r.parts > 1) {
        joined.resize((size_t)r.len);
        rec->copy(r, 0, joined.data());
        payload = joined.data();
      }
      ImageHeader h;
      memcpy(&h, payload, sizeof(h));
      int64_t skip = sizeof(ImageHeader);
      if (h.flag > 0) {
        MXB_CHECK(r.len >= skip + 4 * (int64_t)h.flag, "RecordIO: label array longer than the record");
        for (int k = 0; k < label_width && k < (int)h.flag; ++k)
          memcpy(&s.labels[(size_t)i * label_width + k], payload + skip + 4 * k, 4);
        skip += 4 * (int64_t)h.flag;
      } else {
        s.labels[(size_t)i * label_width] = h.label;
      }
      s.ids[i] = (int64_t)h.id;
      const int64_t nbytes = r.len - skip;
      MXB_CHECK(pos + nbytes <= slot_bytes, "RecordIO loader: staging slot too small for this batch");
      memcpy(s.blob + pos, payload + skip, (size_t)nbytes);
      s.offsets[i] = pos;
      s.lengths[i] = nbytes;
      pos += (nbytes + 15) & ~(int64_t)15;      // 16-byte aligned payloads
    }
    s.offsets[s.count] = pos;
  }

  void worker() {
    for (;;) {
      int64_t b;
      uint64_t ep;
      Slot* s;
      std::vector<int64_t>* ord;
      bool round;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_free.wait(lk, [&] { return stop || (next_claim < num_batches && next_claim < released_upto + depth); });
        if (stop) return;
        b = next_claim++;
        ep = epoch;
        s = &slots[b % depth];
        s->ready = false;
        s->batch_index = b;
        ord = &order;
        round = round_batch;
      }
      std::string err;
      try {
        fill(*s, b, *ord, round);
      } catch (const std::exception& e) {
        err = e.what();
      }
      {
        std::lock_guard<std::mutex> lk(mu);
        if (ep == epoch) {
          s->error = err;
          s->ready = true;
        }
      }
      cv_ready.notify_all();
    }
  }

  void reset(int num_parts, int part_index, bool shuffle, uint64_t seed, bool round) {
    std::unique_lock<std::mutex> lk(mu);
    // let in-flight fills finish: they hold pointers into `order`
    cv_ready.wait(lk, [&] {
      for (int64_t b = released_upto; b < next_claim; ++b)
        if (!slots[b % depth].ready) return false;
      return true;
    });
    MXB_CHECK(num_parts >= 1 && part_index >= 0 && part_index < num_parts, "bad num_parts / part_index");
    const int64_t n = (int64_t)rec->recs.size();
    // contiguous shards like the reference's InputSplit (iter_image_recordio_2.cc:160-166); the remainder goes to the
    // first ranks one record each so every record is visited exactly once per epoch
    const int64_t base = n / num_parts, rem = n % num_parts;
    const int64_t lo = part_index * base + std::min<int64_t>(part_index, rem);
    const int64_t cnt = base + (part_index < rem ? 1 : 0);
    order.resize((size_t)cnt);
    for (int64_t i = 0; i < cnt; ++i) order[(size_t)i] = lo + i;
    if (shuffle) {
      std::mt19937_64 g(seed);
      for (int64_t i = cnt - 1; i > 0; --i) {
        const int64_t j = (int64_t)(g() % (uint64_t)(i + 1));
        std::swap(order[(size_t)i], order[(size_t)j]);
      }
    }
    round_batch = round;
    num_batches = (cnt + batch - 1) / batch;      // round_batch: the last one wraps around and reports `pad`
    next_claim = next_deliver = released_upto = 0;
    for (auto& s : slots) {
      s.ready = false;
      s.batch_index = -1;
    }
    ++epoch;
    lk.unlock();
    cv_free.notify_all();
  }
};

}  // namespace

MXB_API int MXBRecOpen(const char* rec_path, const char* idx_path, void** out) {
  MXB_TRY
  auto* f = new RecFile();
  try {
    f->open(rec_path, idx_path);
  } catch (...) {
    delete f;
    throw;
  }
  *out = f;
  MXB_CATCH
}

MXB_API int MXBRecClose(void* h) {
  MXB_TRY
  delete static_cast<RecFile*>(h);
  MXB_CATCH
}

MXB_API int MXBRecCount(void* h, int64_t* n) {
  MXB_TRY
  *n = (int64_t) static_cast<RecFile*>(h)->recs.size();
  MXB_CATCH
}

// length of record i (logical payload) and its .idx key
MXB_API int MXBRecInfo(void* h, int64_t i, int64_t* len, int64_t* key, int64_t* offset) {
  MXB_TRY
  auto* f = static_cast<RecFile*>(h);
  MXB_CHECK(i >= 0 && i < (int64_t)f->recs.size(), "record index out of range");
  *len = f->recs[(size_t)i].len;
  *key = f->keys[(size_t)i];
  *offset = f->recs[(size_t)i].off;
  MXB_CATCH
}

// copy the payload of record i (recordio framing removed, image header included) into dst[cap]
MXB_API int MXBRecRead(void* h, int64_t i, void* dst, int64_t cap) {
  MXB_TRY
  auto* f = static_cast<RecFile*>(h);
  MXB_CHECK(i >= 0 && i < (int64_t)f->recs.size(), "record index out of range");
  const Rec& r = f->recs[(size_t)i];
  MXB_CHECK(cap >= r.len, "destination too small");
  f->copy(r, 0, static_cast<uint8_t*>(dst));
  MXB_CATCH
}

// upper bound of the packed size of any batch: the `batch` largest records, each rounded up to 16 bytes
MXB_API int MXBRecMaxBatchBytes(void* h, int batch, int64_t* bytes) {
  MXB_TRY
  auto* f = static_cast<RecFile*>(h);
  std::vector<int64_t> lens;
  lens.reserve(f->recs.size());
  for (const Rec& r : f->recs) lens.push_back((r.len + 15) & ~(int64_t)15);
  const size_t k = std::min<size_t>((size_t)batch, lens.size());
  std::partial_sort(lens.begin(), lens.begin() + (long)k, lens.end(), std::greater<int64_t>());
  int64_t s = 0;
  for (size_t i = 0; i < k; ++i) s += lens[i];
  *bytes = s;
  MXB_CATCH
}

// slots[depth]: caller-owned staging buffers (pinned host memory) of slot_bytes each
MXB_API int MXBLoaderCreate(void* rec, int batch, int label_width, int threads, int depth, void** slots, int64_t slot_bytes,
                            void** out) {
  MXB_TRY
  MXB_CHECK(batch > 0 && label_width > 0 && threads > 0 && depth > 0, "bad loader geometry");
  auto* l = new Loader();
  l->rec = static_cast<RecFile*>(rec);
  l->batch = batch;
  l->label_width = label_width;
  l->depth = depth;
  l->slot_bytes = slot_bytes;
  l->slots.resize((size_t)depth);
  for (int i = 0; i < depth; ++i) l->slots[(size_t)i].blob = static_cast<uint8_t*>(slots[i]);
  for (int t = 0; t < threads; ++t) l->workers.emplace_back([l] { l->worker(); });
  *out = l;
  MXB_CATCH
}

// start an epoch: shard = part_index of num_parts (may change between epochs: elastic membership), visiting order
MXB_API int MXBLoaderReset(void* h, int num_parts, int part_index, int shuffle, uint64_t seed, int round_batch) {
  MXB_TRY
  static_cast<Loader*>(h)->reset(num_parts, part_index, shuffle != 0, seed, round_batch != 0);
  MXB_CATCH
}

MXB_API int MXBLoaderNumBatches(void* h, int64_t* n, int64_t* shard_records) {
  MXB_TRY
  auto* l = static_cast<Loader*>(h);
  std::lock_guard<std::mutex> lk(l->mu);
  *n = l->num_batches;
  *shard_records = (int64_t)l->order.size();
  MXB_CATCH
}

// next batch in order.  count == 0: end of epoch.  offsets[batch + 1], lengths[batch], labels[batch * label_width], ids[batch]
MXB_API int MXBLoaderNext(void* h, int* slot, int64_t* offsets, int64_t* lengths, float* labels, int64_t* ids, int* count,
                           int* pad) {
  MXB_TRY
  auto* l = static_cast<Loader*>(h);
  std::unique_lock<std::mutex> lk(l->mu);
  if (l->next_deliver >= l->num_batches) {
    *count = 0;
    *pad = 0;
    *slot = -1;
    return 0;
  }
  const int64_t b = l->next_deliver;
  Slot& s = l->slots[(size_t)(b % l->depth)];
  l->cv_ready.wait(lk, [&] { return s.ready && s.batch_index == b; });
  l->next_deliver++;
  if (!s.error.empty()) throw std::runtime_error(s.error);
  memcpy(offsets, s.offsets.data(), sizeof(int64_t) * (size_t)(s.count + 1));
  memcpy(lengths, s.lengths.data(), sizeof(int64_t) * (size_t)s.count);
  memcpy(labels, s.labels.data(), sizeof(float) * s.labels.size());
  memcpy(ids, s.ids.data(), sizeof(int64_t) * (size_t)s.count);
  *count = s.count;
  *pad = s.pad;
  *slot = (int)(b % l->depth);
  MXB_CATCH
}

// the consumer is done with the OLDEST outstanding batch (batches are released in delivery order)
MXB_API int MXBLoaderRelease(void* h) {
  MXB_TRY
  auto* l = static_cast<Loader*>(h);
  {
    std::lock_guard<std::mutex> lk(l->mu);
    MXB_CHECK(l->released_upto < l->next_deliver, "release without a delivered batch");
    l->released_upto++;
  }
  l->cv_free.notify_all();
  MXB_CATCH
}

MXB_API int MXBLoaderFree(void* h) {
  MXB_TRY
  auto* l = static_cast<Loader*>(h);
  {
    std::lock_guard<std::mutex> lk(l->mu);
    l->stop = true;
  }
  l->cv_free.notify_all();
  for (auto& t : l->workers) t.join();
  delete l;
  MXB_CATCH
}

// Frame size of n JPEG streams without decoding them: scan the marker segments up to the first start-of-frame.
// widths[i] = heights[i] = 0 marks a payload that is not a JPEG (PNG, raw): the caller routes it to the CPU decoder.
MXB_API int MXBJpegDims(const void* blob, const int64_t* offsets, const int64_t* lengths, int n, int* widths, int* heights) {
  MXB_TRY
  for (int i = 0; i < n; ++i) {
    const uint8_t* p = static_cast<const uint8_t*>(blob) + offsets[i];
    const int64_t len = lengths[i];
    widths[i] = heights[i] = 0;
    if (len < 4 || p[0] != 0xFF || p[1] != 0xD8) continue;
    int64_t q = 2;
    while (q + 4 <= len) {
      if (p[q] != 0xFF) { ++q; continue; }
      const uint8_t m = p[q + 1];
      if (m == 0xFF) { ++q; continue; }                       // fill byte
      if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) { q += 2; continue; }   // markers without a length
      const int64_t seg = ((int64_t)p[q + 2] << 8) | p[q + 3];
      const bool sof = m >= 0xC0 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC;
      if (sof) {
        if (q + 9 <= len) {
          heights[i] = ((int)p[q + 5] << 8) | p[q + 6];
          widths[i] = ((int)p[q + 7] << 8) | p[q + 8];
        }
        break;
      }
      if (m == 0xDA || m == 0xD9) break;                      // start of scan / end of image before any frame header
      q += 2 + seg;
    }
  }
  MXB_CATCH
}
