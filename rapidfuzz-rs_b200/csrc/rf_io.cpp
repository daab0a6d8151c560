// rf_io.cpp -- the step BEFORE the scoring path: getting candidates into the packed
// CSR form the kernels read.  The reference consumes one arbitrary iterator per call
// (levenshtein.rs:1750-1762); real callers hold a Vec<String>.  Host code only (OpenMP + mmap):
//   rf_pack_u8              array of (pointer, length) strings -> offsets[n+1] + chars[total], parallel copy
//   rf_corpus_file_write    CSR corpus -> one file that can be mapped back without parsing
//   rf_corpus_file_open     mmap a corpus file; accessors hand out pointers into the mapping, which can be fed to
//                           rf_corpus_create_* / rf_batch_stream_* directly
// File layout (little endian, every section 64-byte aligned):
//   [0,64)   header: magic "RFCORP01", u32 version = 1, u32 elem_size = 1, u64 n, u64 total, u32 offset_width (4|8)
//   [64,..)  offsets[n+1] as u32 (total < 2^32 - 16) or u64
//   [..,..)  chars[total]
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <new>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../../include/rfgpu.h"

namespace {
constexpr char kMagic[8] = {'R', 'F', 'C', 'O', 'R', 'P', '0', '1'};
struct FileHeader {
  char magic[8];
  uint32_t version, elem_size;
  uint64_t n, total;
  uint32_t offset_width, pad0;
  uint8_t reserved[24];
};
static_assert(sizeof(FileHeader) == 64, "header is one 64-byte section");
inline uint64_t align64(uint64_t x) { return (x + 63) & ~63ull; }

}  // namespace
extern "C" void rf__set_last_error(const char* msg);  // rf_api.cu: the thread-local text behind rf_last_error()
namespace {
rf_status io_fail(rf_status s, const std::string& msg) {
  rf__set_last_error(msg.c_str());
  return s;
}
}  // namespace

struct rf_corpus_file {
  void* map = nullptr;
  uint64_t map_bytes = 0;
  uint64_t n = 0, total = 0;
  uint32_t offset_width = 0;
  const void* offsets = nullptr;
  const uint8_t* chars = nullptr;
};

extern "C" {

rf_status rf_pack_u8(const uint8_t* const* strings, const uint64_t* lengths, uint64_t n, uint64_t* offsets_out,
                     uint8_t* chars_out, int nthreads) {
  if (!offsets_out) return io_fail(RF_ERR_INVALID_ARG, "offsets_out is NULL");
  if (n && !lengths) return io_fail(RF_ERR_INVALID_ARG, "lengths is NULL");
  uint64_t run = 0;
  for (uint64_t i = 0; i < n; ++i) {  // prefix sum (sequential: memory-bound and tiny next to the copy)
    offsets_out[i] = run;
    run += lengths[i];
  }
  offsets_out[n] = run;
  if (!chars_out) return RF_OK;  // sizing call
  if (n && !strings) return io_fail(RF_ERR_INVALID_ARG, "strings is NULL");
#ifdef _OPENMP
  const int threads = nthreads > 0 ? nthreads : omp_get_max_threads();
#pragma omp parallel for schedule(static, 4096) num_threads(threads)
#endif
  for (int64_t i = 0; i < (int64_t)n; ++i)
    if (lengths[i]) memcpy(chars_out + offsets_out[i], strings[i], lengths[i]);
  (void)nthreads;
  return RF_OK;
}

// 6-bit packing of a byte corpus with at most 64 distinct symbols (ASCII alphanumerics: 62): 4 characters -> 3 bytes, codes
// in ascending byte order.  The packed stream crosses PCIe 25 % smaller (rf_batch_stream_*_packed6 unpacks it on the device).
uint64_t rf_pack6_size(uint64_t total_chars) { return (total_chars + 3) / 4 * 3 + 64; }

rf_status rf_pack6_u8(const uint8_t* chars, uint64_t total, uint8_t* packed_out, uint8_t* dict_out, int nthreads) {
  if (!dict_out || !packed_out) return io_fail(RF_ERR_INVALID_ARG, "NULL output");
  if (total && !chars) return io_fail(RF_ERR_INVALID_ARG, "chars is NULL");
#ifdef _OPENMP
  const int threads = nthreads > 0 ? nthreads : omp_get_max_threads();
#else
  const int threads = 1;
#endif
  (void)threads;
  uint64_t seen[4] = {0, 0, 0, 0};
#pragma omp parallel num_threads(threads)
  {
    uint64_t mine[4] = {0, 0, 0, 0};
#pragma omp for schedule(static) nowait
    for (int64_t i = 0; i < (int64_t)total; ++i) mine[chars[i] >> 6] |= 1ull << (chars[i] & 63);
#pragma omp critical
    for (int k = 0; k < 4; ++k) seen[k] |= mine[k];
  }
  uint8_t code_of[256];
  memset(code_of, 0, sizeof(code_of));
  memset(dict_out, 0, 64);
  uint32_t d = 0;
  for (int v = 0; v < 256; ++v) {
    if (!((seen[v >> 6] >> (v & 63)) & 1)) continue;
    if (d == 64) return io_fail(RF_ERR_UNSUPPORTED, "more than 64 distinct symbols: the corpus cannot be packed to 6 bits");
    code_of[v] = (uint8_t)d;
    dict_out[d++] = (uint8_t)v;
  }
  const uint64_t quads = (total + 3) / 4;
#pragma omp parallel for schedule(static) num_threads(threads)
  for (int64_t qd = 0; qd < (int64_t)quads; ++qd) {
    uint32_t c[4];
    for (int j = 0; j < 4; ++j) c[j] = ((uint64_t)qd * 4 + j < total) ? code_of[chars[(uint64_t)qd * 4 + j]] : 0u;
    const uint32_t w = c[0] | (c[1] << 6) | (c[2] << 12) | (c[3] << 18);
    packed_out[qd * 3 + 0] = (uint8_t)w;
    packed_out[qd * 3 + 1] = (uint8_t)(w >> 8);
    packed_out[qd * 3 + 2] = (uint8_t)(w >> 16);
  }
  memset(packed_out + quads * 3, 0, 64);
  return RF_OK;
}

rf_status rf_corpus_file_write(const char* path, const uint8_t* chars, const uint64_t* offsets, uint64_t n) {
  if (!path || !offsets) return io_fail(RF_ERR_INVALID_ARG, "NULL argument");
  if (offsets[0] != 0) return io_fail(RF_ERR_INVALID_ARG, "offsets[0] must be 0");
  const uint64_t total = offsets[n];
  if (total && !chars) return io_fail(RF_ERR_INVALID_ARG, "chars is NULL");
  for (uint64_t i = 0; i < n; ++i)
    if (offsets[i + 1] < offsets[i]) return io_fail(RF_ERR_INVALID_ARG, "offsets must be non-decreasing");
  FileHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, kMagic, 8);
  h.version = 1;
  h.elem_size = 1;
  h.n = n;
  h.total = total;
  h.offset_width = total < 0xFFFFFFF0ull ? 4 : 8;
  FILE* f = fopen(path, "wb");
  if (!f) return io_fail(RF_ERR_INVALID_ARG, std::string("cannot create ") + path + ": " + strerror(errno));
  bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
  const uint64_t off_bytes = (n + 1) * h.offset_width;
  if (ok && h.offset_width == 8) ok = fwrite(offsets, 8, n + 1, f) == n + 1;
  if (ok && h.offset_width == 4) {
    std::vector<uint32_t> buf(1 << 20);
    for (uint64_t i = 0; ok && i <= n;) {
      const uint64_t m = (n + 1 - i) < buf.size() ? (n + 1 - i) : buf.size();
      for (uint64_t j = 0; j < m; ++j) buf[j] = (uint32_t)offsets[i + j];
      ok = fwrite(buf.data(), 4, m, f) == m;
      i += m;
    }
  }
  static const uint8_t zeros[64] = {0};
  const uint64_t pad = align64(sizeof(h) + off_bytes) - (sizeof(h) + off_bytes);
  if (ok && pad) ok = fwrite(zeros, 1, pad, f) == pad;
  if (ok && total) ok = fwrite(chars, 1, total, f) == total;
  if (fclose(f) != 0) ok = false;
  if (!ok) return io_fail(RF_ERR_INVALID_ARG, std::string("write failed: ") + path);
  return RF_OK;
}

rf_status rf_corpus_file_open(const char* path, rf_corpus_file** out) {
  if (!out) return io_fail(RF_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (!path) return io_fail(RF_ERR_INVALID_ARG, "path is NULL");
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return io_fail(RF_ERR_INVALID_ARG, std::string("cannot open ") + path + ": " + strerror(errno));
  struct stat st;
  if (fstat(fd, &st) != 0 || (uint64_t)st.st_size < sizeof(FileHeader)) {
    close(fd);
    return io_fail(RF_ERR_INVALID_ARG, "not a corpus file (too short)");
  }
  void* map = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (map == MAP_FAILED) return io_fail(RF_ERR_OOM, std::string("mmap failed: ") + strerror(errno));
  const FileHeader* h = (const FileHeader*)map;
  const bool hdr_ok = memcmp(h->magic, kMagic, 8) == 0 && h->version == 1 && h->elem_size == 1 &&
                      (h->offset_width == 4 || h->offset_width == 8) && h->n < 0xFFFFFFFFull &&
                      h->total <= (uint64_t)st.st_size;  // also keeps the size arithmetic below from wrapping
  uint64_t need = 0;
  if (hdr_ok) need = align64(sizeof(FileHeader) + (h->n + 1) * (uint64_t)h->offset_width) + h->total;
  if (!hdr_ok || need > (uint64_t)st.st_size) {
    munmap(map, (size_t)st.st_size);
    return io_fail(RF_ERR_INVALID_ARG, "not a corpus file (bad header or truncated)");
  }
  rf_corpus_file* cf = new (std::nothrow) rf_corpus_file();
  if (!cf) {
    munmap(map, (size_t)st.st_size);
    return io_fail(RF_ERR_OOM, "host allocation failed");
  }
  cf->map = map;
  cf->map_bytes = (uint64_t)st.st_size;
  cf->n = h->n;
  cf->total = h->total;
  cf->offset_width = h->offset_width;
  cf->offsets = (const uint8_t*)map + sizeof(FileHeader);
  cf->chars = (const uint8_t*)map + align64(sizeof(FileHeader) + (h->n + 1) * (uint64_t)h->offset_width);
  // cheap integrity check of the CSR ends
  const uint64_t first = h->offset_width == 4 ? ((const uint32_t*)cf->offsets)[0] : ((const uint64_t*)cf->offsets)[0];
  const uint64_t last = h->offset_width == 4 ? ((const uint32_t*)cf->offsets)[h->n] : ((const uint64_t*)cf->offsets)[h->n];
  if (first != 0 || last != h->total) {
    rf_corpus_file_close(cf);
    return io_fail(RF_ERR_INVALID_ARG, "corpus file: offsets do not match the header");
  }
  // full integrity check of the CSR index (parallel; one pass over the offsets): the mapping is handed to the scan
  // kernels as is by rf_batch_stream_*, and a decreasing or out-of-range start would send them out of bounds
  {
    const uint64_t n = h->n, total = h->total;
    int bad = 0;
    if (h->offset_width == 4) {
      const uint32_t* o = (const uint32_t*)cf->offsets;
#pragma omp parallel for schedule(static) reduction(| : bad)
      for (int64_t i = 0; i < (int64_t)n; ++i) bad |= (o[i + 1] < o[i]) | ((uint64_t)o[i + 1] > total);
    } else {
      const uint64_t* o = (const uint64_t*)cf->offsets;
#pragma omp parallel for schedule(static) reduction(| : bad)
      for (int64_t i = 0; i < (int64_t)n; ++i) bad |= (o[i + 1] < o[i]) | (o[i + 1] > total);
    }
    if (bad) {
      rf_corpus_file_close(cf);
      return io_fail(RF_ERR_INVALID_ARG, "corpus file: offsets are not non-decreasing CSR starts");
    }
  }
  madvise(map, (size_t)st.st_size, MADV_SEQUENTIAL);
  *out = cf;
  return RF_OK;
}

rf_status rf_corpus_file_close(rf_corpus_file* f) {
  if (!f) return RF_OK;
  if (f->map) munmap(f->map, (size_t)f->map_bytes);
  delete f;
  return RF_OK;
}
uint64_t rf_corpus_file_size(const rf_corpus_file* f) { return f ? f->n : 0; }
uint64_t rf_corpus_file_total_chars(const rf_corpus_file* f) { return f ? f->total : 0; }
uint32_t rf_corpus_file_offset_width(const rf_corpus_file* f) { return f ? f->offset_width : 0; }
const void* rf_corpus_file_offsets(const rf_corpus_file* f) { return f ? f->offsets : nullptr; }
const uint8_t* rf_corpus_file_chars(const rf_corpus_file* f) { return f ? f->chars : nullptr; }

rf_status rf_corpus_create_from_file(const char* path, int device, rf_corpus** out) {
  if (!out) return io_fail(RF_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  rf_corpus_file* f = nullptr;
  rf_status s = rf_corpus_file_open(path, &f);
  if (s != RF_OK) return s;
  if (f->offset_width == 4) s = rf_corpus_create_u8_off32(f->chars, (const uint32_t*)f->offsets, f->n, device, out);
  else s = rf_corpus_create_u8(f->chars, (const uint64_t*)f->offsets, f->n, device, out);
  rf_corpus_file_close(f);
  return s;
}

}  // extern "C"
