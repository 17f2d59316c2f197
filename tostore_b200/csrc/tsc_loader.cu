// tsc_loader.cu — cold start: stream an on-disk ToStore NGH vector index into a GPU
// index. Host code; the bytes of the reference's partition files
// go to tsc_index_append_pages / tsc_index_apply_graph_pages untouched (magic / CRC /
// dims validation and decode happen on the GPU, tsc_ingest.cuh).
//
// A reader thread walks the partition files and fills two pinned staging buffers while
// the calling thread feeds the previous buffer to the GPU, so disk reads overlap the
// host->device copies and the validate / decode kernels.
//
// Layout facts restated from the reference (paths relative to /root/reference/lib/src):
//   <index>/ngh/meta.json, top-level keys            model/ngh_index_meta.dart:410-446
//   <index>/ngh/<rawvec|graph>/dir_<p ~/ 500>/p<p>.ngh   core/path_manager.dart:317-324,
//                                                    handler/common.dart:43
//   page 0 of every file is a per-file meta page     core/ngh_page.dart:29-98
//   P = maxPartitionFileSize ~/ nghPageSize data pages per file; logical page l lives in
//   file l ~/ P at local page 1 + l % P              model/ngh_index_meta.dart:178, :451-490
//   rows per raw page  = (pageSize-20-8-64) ~/ (dims*bpe)          core/ngh_page.dart:575-579
//   slots per graph page = (pageSize-20-4-64) ~/ (2+4*maxDegree)   core/ngh_page.dart:559-566
// Extents come from nextNodeId, never from the *PartitionCount fields (the reference does
// not maintain them). Missing or short files contribute the pages that
// exist (readers in the reference degrade to empty pages, ngh_partition_manager.dart:262-287).
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <functional>
#include <map>
#include <string>
#include <thread>

#include "tsc_index.h"

using namespace tsc;

namespace {

constexpr uint32_t kMaxEntriesPerDir = 500;           // handler/common.dart:43
constexpr uint64_t kChunkBytes = 64ull << 20;         // one staging buffer

// ---- meta.json: top-level keys only (nested objects such as nodeIdToPkMeta repeat names)
bool json_top_level(const std::string &t, std::map<std::string, std::string> *out) {
  size_t i = t.find('{');
  if (i == std::string::npos) return false;
  i++;
  auto skip_ws = [&]() {
    while (i < t.size() && (t[i] == ' ' || t[i] == '\n' || t[i] == '\r' || t[i] == '\t')) i++;
  };
  auto parse_string = [&](std::string *s) -> bool {   // t[i] == '"'
    i++;
    while (i < t.size() && t[i] != '"') {
      if (t[i] == '\\' && i + 1 < t.size()) {
        if (s) s->push_back(t[i + 1]);
        i += 2;
      } else {
        if (s) s->push_back(t[i]);
        i++;
      }
    }
    if (i >= t.size()) return false;
    i++;
    return true;
  };
  for (;;) {
    skip_ws();
    if (i >= t.size()) return false;
    if (t[i] == '}') return true;
    if (t[i] == ',') {
      i++;
      continue;
    }
    if (t[i] != '"') return false;
    std::string key, val;
    if (!parse_string(&key)) return false;
    skip_ws();
    if (i >= t.size() || t[i] != ':') return false;
    i++;
    skip_ws();
    if (i >= t.size()) return false;
    if (t[i] == '"') {
      if (!parse_string(&val)) return false;
    } else if (t[i] == '{' || t[i] == '[') {
      int depth = 0;
      while (i < t.size()) {
        if (t[i] == '"') {
          if (!parse_string(nullptr)) return false;
          continue;
        }
        if (t[i] == '{' || t[i] == '[') depth++;
        if (t[i] == '}' || t[i] == ']') depth--;
        i++;
        if (depth == 0) break;
      }
      if (depth != 0) return false;
      val = "{}";
    } else {
      while (i < t.size() && t[i] != ',' && t[i] != '}' && t[i] != ' ' && t[i] != '\n' &&
             t[i] != '\r' && t[i] != '\t')
        val.push_back(t[i++]);
    }
    (*out)[key] = val;
  }
}

int32_t read_meta(const char *index_dir, tsc_ngh_info *m) {
  const std::string path = std::string(index_dir) + "/ngh/meta.json";
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) {
    set_error("ngh meta: cannot open %s: %s", path.c_str(), strerror(errno));
    return TSC_ERR_BAD_ARG;
  }
  std::string text;
  char buf[4096];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
  fclose(f);
  std::map<std::string, std::string> kv;
  if (!json_top_level(text, &kv)) {
    set_error("ngh meta: %s is not a JSON object", path.c_str());
    return TSC_ERR_PAGE;
  }
  auto num = [&](const char *key, uint64_t dflt) -> uint64_t {
    auto it = kv.find(key);
    if (it == kv.end() || it->second.empty() || it->second == "null") return dflt;
    return (uint64_t)strtod(it->second.c_str(), nullptr);   // `(json[..] as num?)?.toInt()`
  };
  auto str = [&](const char *key) -> std::string {
    auto it = kv.find(key);
    return it == kv.end() ? std::string() : it->second;
  };
  if (kv.find("dimensions") == kv.end()) {
    set_error("ngh meta: %s has no \"dimensions\"", path.c_str());
    return TSC_ERR_PAGE;
  }
  m->dims = (uint32_t)num("dimensions", 0);
  const std::string metric = str("distanceMetric"), prec = str("precision");
  // _parseDistanceMetric / _parsePrecision defaults: cosine, float32 (ngh_index_meta.dart:494-516)
  m->metric = metric == "l2" ? TSC_METRIC_L2
                             : (metric == "innerProduct" ? TSC_METRIC_INNER_PRODUCT : TSC_METRIC_COSINE);
  m->precision = prec == "float64" ? TSC_SRC_F64 : (prec == "int8" ? TSC_SRC_I8 : TSC_SRC_F32);
  m->next_node_id = num("nextNodeId", 0);
  m->page_size = (uint32_t)num("nghPageSize", 16 * 1024);
  m->max_partition_file_size = num("maxPartitionFileSize", 16ull * 1024 * 1024);
  m->max_degree = (uint32_t)num("maxDegree", 64);
  if (m->dims == 0 || m->page_size < 128 || m->max_partition_file_size < m->page_size) {
    set_error("ngh meta: implausible dimensions=%u nghPageSize=%u maxPartitionFileSize=%llu",
              m->dims, m->page_size, (unsigned long long)m->max_partition_file_size);
    return TSC_ERR_PAGE;
  }
  return TSC_OK;
}

uint32_t rows_per_raw_page(const tsc_ngh_info &m) {
  const uint32_t bpe = m.precision == TSC_SRC_F64 ? 8 : (m.precision == TSC_SRC_I8 ? 1 : 4);
  const int64_t usable = (int64_t)m.page_size - 20 - 8 - 64;
  return usable > 0 ? (uint32_t)(usable / ((int64_t)m.dims * bpe)) : 0;
}
uint32_t slots_per_graph_page(const tsc_ngh_info &m) {
  const int64_t usable = (int64_t)m.page_size - 20 - 4 - 64;
  return usable > 0 ? (uint32_t)(usable / (2 + (int64_t)m.max_degree * 4)) : 0;
}

std::string partition_path(const char *index_dir, const char *category, uint64_t part) {
  char tail[96];
  snprintf(tail, sizeof tail, "/ngh/%s/dir_%llu/p%llu.ngh", category,
           (unsigned long long)(part / kMaxEntriesPerDir), (unsigned long long)part);
  return std::string(index_dir) + tail;
}

struct Chunk {
  uint64_t first_page = 0, n_pages = 0;
  bool last = false;
};
using Sink = std::function<int32_t(const Chunk &, const uint8_t *)>;

// Walk the data pages that hold node ids [node_lo, node_hi) of one category and hand them
// to `sink` in chunks of whole, logically consecutive pages (never more than chunk_pages,
// never across a partition file). Reading runs one chunk ahead of the sink on its own
// thread, into two buffers.
int32_t walk_pages(const char *index_dir, const tsc_ngh_info &m, const char *category,
                   uint32_t per_page, uint64_t node_lo, uint64_t node_hi, uint64_t chunk_pages,
                   uint8_t *buf0, uint8_t *buf1, const Sink &sink, tsc_ngh_info *stats) {
  if (per_page == 0 || node_hi <= node_lo) return TSC_OK;
  const uint64_t ps = m.page_size;
  const uint64_t P = m.max_partition_file_size / ps;
  const uint64_t lp_lo = node_lo / per_page, lp_hi = (node_hi + per_page - 1) / per_page;

  std::mutex mu;
  std::condition_variable cv;
  Chunk ready[2];
  bool full[2] = {false, false};
  bool abort = false;
  uint8_t *bufs[2] = {buf0, buf1};
  uint64_t files = 0, pages = 0;

  std::thread reader([&]() {
    int slot = 0;
    auto publish = [&](const Chunk &c) -> bool {
      std::unique_lock<std::mutex> lk(mu);
      ready[slot] = c;
      full[slot] = true;
      cv.notify_all();
      slot ^= 1;
      cv.wait(lk, [&] { return !full[slot] || abort; });   // next buffer must be free
      return !abort;
    };
    bool alive = true;                       // nothing published yet: both buffers are free
    for (uint64_t part = lp_lo / P; alive && part * P < lp_hi; part++) {
      const uint64_t a = part * P > lp_lo ? part * P : lp_lo;
      const uint64_t b = (part + 1) * P < lp_hi ? (part + 1) * P : lp_hi;
      const std::string path = partition_path(index_dir, category, part);
      const int fd = open(path.c_str(), O_RDONLY);
      if (fd < 0) continue;                      // missing file: its pages are simply absent
      files++;
      for (uint64_t lp = a; alive && lp < b;) {
        const uint64_t want = b - lp < chunk_pages ? b - lp : chunk_pages;
        const off_t off = (off_t)((1 + lp % P) * ps);   // page 0 = per-file meta page
        uint64_t got = 0;
        while (got < want * ps) {
          const ssize_t r = pread(fd, bufs[slot] + got, want * ps - got, off + (off_t)got);
          if (r <= 0) break;
          got += (uint64_t)r;
        }
        const uint64_t whole = got / ps;
        if (whole) {
          pages += whole;
          Chunk c;
          c.first_page = lp;
          c.n_pages = whole;
          alive = publish(c);
        }
        if (whole < want) break;                 // short file: what exists has been delivered
        lp += whole;
      }
      close(fd);
    }
    std::unique_lock<std::mutex> lk(mu);
    Chunk end;
    end.last = true;
    cv.wait(lk, [&] { return !full[slot] || abort; });
    ready[slot] = end;
    full[slot] = true;
    cv.notify_all();
  });

  int32_t rc = TSC_OK;
  for (int slot = 0;; slot ^= 1) {
    Chunk c;
    {
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] { return full[slot]; });
      c = ready[slot];
    }
    if (!c.last && rc == TSC_OK) rc = sink(c, bufs[slot]);
    {
      std::unique_lock<std::mutex> lk(mu);
      full[slot] = false;
      if (rc != TSC_OK) abort = true;
      cv.notify_all();
    }
    if (c.last) break;
  }
  reader.join();
  if (stats) {
    stats->files_read += files;
    stats->pages_read += pages;
    stats->bytes_read += pages * ps;
  }
  return rc;
}

uint32_t crc32_ieee(const uint8_t *p, size_t n) {
  static uint32_t tab[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; i++) {
      uint32_t c = i;
      for (int k = 0; k < 8; k++) c = (c & 1u) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
      tab[i] = c;
    }
    init = true;
  }
  uint32_t c = 0xFFFFFFFFu;
  for (size_t i = 0; i < n; i++) c = tab[(c ^ p[i]) & 0xFFu] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

}  // namespace

extern "C" {

int32_t tsc_ngh_read_meta(const char *index_dir, tsc_ngh_info *out) {
  if (!index_dir || !out || out->struct_size != sizeof(tsc_ngh_info)) {
    set_error("ngh_read_meta: NULL argument or struct_size mismatch");
    return TSC_ERR_BAD_ARG;
  }
  memset((uint8_t *)out + 4, 0, sizeof(*out) - 4);
  return read_meta(index_dir, out);
}

int32_t tsc_index_load_ngh(uint64_t handle, const char *index_dir, uint32_t flags,
                           tsc_ngh_info *out) {
  TSC_API_TRY
  if (GroupRef g = lookup_group(handle)) {
    std::lock_guard<std::mutex> glk(g->mu);
    return grp_load_ngh(*g, index_dir, flags, out);
  }
  IndexRef ref = lookup_index(handle);
  if (!ref) return TSC_ERR_BAD_HANDLE;
  if (ref->host_only) {
    set_error("host-only self-test handle: no device entry point works on it");
    return TSC_ERR_UNSUPPORTED;
  }
  return ix_load_ngh(ref.get(), index_dir, flags, out);
  TSC_API_CATCH
}

}  // extern "C"

namespace tsc {
int32_t ix_load_ngh(Index *ix, const char *index_dir, uint32_t flags, tsc_ngh_info *out) {
  tsc_ngh_info m;
  memset(&m, 0, sizeof m);
  m.struct_size = sizeof m;
  if (!index_dir || (out && out->struct_size != sizeof(tsc_ngh_info))) {
    set_error("index_load_ngh: NULL directory or struct_size mismatch");
    return TSC_ERR_BAD_ARG;
  }
  int32_t rc = read_meta(index_dir, &m);
  if (rc != TSC_OK) return rc;
  if (m.dims != ix->desc.dims || m.precision != ix->desc.src_precision) {
    set_error("index_load_ngh: index on disk is dims=%u precision=%u, GPU index dims=%u "
              "src_precision=%u", m.dims, m.precision, ix->desc.dims, ix->desc.src_precision);
    return TSC_ERR_BAD_DIMS;
  }
  const auto t0 = std::chrono::steady_clock::now();
  // this shard's share of the node ids
  const uint64_t lo = ix->desc.first_node_id;
  uint64_t hi = lo + ix->capacity;
  if (hi > m.next_node_id) hi = m.next_node_id;
  const uint64_t chunk_pages = kChunkBytes / m.page_size ? kChunkBytes / m.page_size : 1;
  uint8_t *bufs[2] = {nullptr, nullptr};
  cudaSetDevice(ix->device);
  for (int i = 0; i < 2; i++)
    if (cudaHostAlloc((void **)&bufs[i], chunk_pages * m.page_size, cudaHostAllocDefault) !=
        cudaSuccess) {
      cudaGetLastError();
      set_error("index_load_ngh: cannot allocate pinned staging buffers");
      for (int j = 0; j < i; j++) cudaFreeHost(bufs[j]);
      return TSC_ERR_OOM;
    }
  const uint32_t page_size = m.page_size;
  const uint64_t live = m.next_node_id;
  rc = walk_pages(index_dir, m, "rawvec", rows_per_raw_page(m), lo, hi, chunk_pages, bufs[0],
                  bufs[1],
                  [&](const Chunk &c, const uint8_t *data) {
                    return ix_append_pages(ix, c.first_page, data, c.n_pages, page_size, live);
                  },
                  &m);
  if (rc == TSC_OK && (flags & TSC_LOAD_TOMBSTONES))
    rc = walk_pages(index_dir, m, "graph", slots_per_graph_page(m), lo, hi, chunk_pages, bufs[0],
                    bufs[1],
                    [&](const Chunk &c, const uint8_t *data) {
                      return ix_apply_graph_pages(ix, c.first_page, data, c.n_pages, page_size);
                    },
                    &m);
  cudaFreeHost(bufs[0]);
  cudaFreeHost(bufs[1]);
  m.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (out) *out = m;
  return rc;
}
}  // namespace tsc

extern "C" {

// Self-test hook (no GPU): the directory walk / chunking / double-buffered reader of
// tsc_index_load_ngh with a host sink that records (first logical page, pages, CRC-32 of the
// chunk's bytes). category 0 = rawvec, 1 = graph.
int32_t tsc_selftest_ngh_walk(const char *index_dir, uint32_t category, uint64_t node_lo,
                              uint64_t node_hi, uint32_t chunk_pages, uint64_t *out_first_page,
                              uint64_t *out_n_pages, uint32_t *out_crc, uint32_t max_chunks,
                              uint32_t *out_n_chunks) {
  if (!index_dir || !out_first_page || !out_n_pages || !out_crc || !out_n_chunks ||
      chunk_pages == 0 || category > 1) {
    set_error("selftest_ngh_walk: bad argument");
    return TSC_ERR_BAD_ARG;
  }
  tsc_ngh_info m;
  memset(&m, 0, sizeof m);
  m.struct_size = sizeof m;
  int32_t rc = read_meta(index_dir, &m);
  if (rc != TSC_OK) return rc;
  if (node_hi > m.next_node_id) node_hi = m.next_node_id;
  std::vector<uint8_t> b0((size_t)chunk_pages * m.page_size), b1((size_t)chunk_pages * m.page_size);
  uint32_t n = 0;
  const uint32_t ps = m.page_size;
  rc = walk_pages(index_dir, m, category == 0 ? "rawvec" : "graph",
                  category == 0 ? rows_per_raw_page(m) : slots_per_graph_page(m), node_lo, node_hi,
                  chunk_pages, b0.data(), b1.data(),
                  [&](const Chunk &c, const uint8_t *data) -> int32_t {
                    if (n >= max_chunks) {
                      set_error("selftest_ngh_walk: more than %u chunks", max_chunks);
                      return TSC_ERR_OOM;
                    }
                    out_first_page[n] = c.first_page;
                    out_n_pages[n] = c.n_pages;
                    out_crc[n] = crc32_ieee(data, (size_t)c.n_pages * ps);
                    n++;
                    return TSC_OK;
                  },
                  nullptr);
  *out_n_chunks = n;
  return rc;
}

}  // extern "C"
