// Frame ingest: PNG (or pre-decoded PPM / PGM / .npy) files -> uint8 [n, H, W, 3] frames in caller memory (normally the pinned
// host buffers BetaposeEngine.run_stream uploads from), decoded by a pool of host threads while the GPU works on the
// previous batch.  Stands in for cv2.imread / PIL.Image.open in ImageLoader.getitem_yolo (dataloader.py:150-179) and
// prep_image (yolo/preprocess.py:34-46), which decode every frame twice on one Python thread.
//
// Host-only code: nothing here touches CUDA, so it also runs (and is tested) on a box without a GPU.  Container parsing,
// inflate (inflate_fast.h: zlib's own inflate() was 90 % of the decode time), un-filtering and sample conversion are done
// here; zlib only supplies crc32().  Conversion rules are the
// ones both reference decoders share for an 8-bit 3-channel result: alpha dropped (no compositing), grey replicated,
// palette expanded, 1/2/4-bit grey scaled to 0..255, 16-bit samples reduced to their high byte, gamma ignored.
// Interlaced (Adam7) files are reported as BP_ERR_UNSUPPORTED (the Python layer hands those to Pillow).
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "betapose_b200.h"
#include "inflate_fast.h"  // also brings <emmintrin.h> when SSE2 is there

int bp_fail(int code, const char* msg);

namespace {

struct PngHeader {
  uint32_t w = 0, h = 0;
  int depth = 0, ctype = 0, interlace = 0;
  int channels() const { return ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : 4; }
};

inline uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

const uint8_t kSig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};

struct Scratch {  // per decoding thread, reused between files
  std::vector<uint8_t> file, comp, raw;
  bp_inflate::Tables tables;
};

int parse_header(const uint8_t* png, size_t len, PngHeader* hd, std::string* err) {
  if (len < 8 + 25 || memcmp(png, kSig, 8) != 0) {
    *err = "not a PNG stream";
    return BP_ERR_UNSUPPORTED;
  }
  if (be32(png + 8) != 13 || memcmp(png + 12, "IHDR", 4) != 0) {
    *err = "PNG: first chunk is not IHDR";
    return BP_ERR_INVALID;
  }
  const uint8_t* d = png + 16;
  hd->w = be32(d);
  hd->h = be32(d + 4);
  hd->depth = d[8];
  hd->ctype = d[9];
  hd->interlace = d[12];
  const int dp = hd->depth;
  bool ok = false;
  switch (hd->ctype) {
    case 0: ok = dp == 1 || dp == 2 || dp == 4 || dp == 8 || dp == 16; break;
    case 3: ok = dp == 1 || dp == 2 || dp == 4 || dp == 8; break;
    case 2: case 4: case 6: ok = dp == 8 || dp == 16; break;
    default: break;
  }
  if (!ok || d[10] != 0 || d[11] != 0 || hd->interlace > 1 || hd->w == 0 || hd->h == 0 || hd->w > (1u << 16) || hd->h > (1u << 16)) {
    *err = "PNG: bad IHDR";
    return BP_ERR_INVALID;
  }
  return BP_OK;
}

inline int paeth(int a, int b, int c) {
  const int p = a + b - c;
  const int pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

#if defined(__SSE2__)
// Average and Paeth depend on the pixel just reconstructed, so they run pixel by pixel -- but with the 3 or 4 samples of
// a pixel side by side in 16-bit lanes and without branches: the scalar Paeth predictor mispredicts on almost every
// sample of a camera image (6.8 ms per 640x480 frame against 0.9 ms this way).  Loads fetch 4 bytes even for 3-byte
// pixels (the byte after a row is the next row's filter byte or the buffer's slack); stores write exactly BPP bytes.
inline __m128i px_load(const uint8_t* p) {
  uint32_t v;
  memcpy(&v, p, 4);
  return _mm_unpacklo_epi8(_mm_cvtsi32_si128(int(v)), _mm_setzero_si128());
}
template <int BPP>
inline void px_store(uint8_t* p, __m128i v16) {
  const uint32_t v = uint32_t(_mm_cvtsi128_si32(_mm_packus_epi16(v16, v16)));
  memcpy(p, &v, BPP);
}
inline __m128i abs16(__m128i x) { return _mm_max_epi16(x, _mm_sub_epi16(_mm_setzero_si128(), x)); }

template <int BPP>
void paeth_row_simd(uint8_t* cur, const uint8_t* prev, size_t n) {
  const __m128i low8 = _mm_set1_epi16(0xFF);
  __m128i a = _mm_setzero_si128(), c = _mm_setzero_si128();
  for (size_t i = 0; i + BPP <= n; i += BPP) {
    const __m128i b = px_load(prev + i), x = px_load(cur + i);
    const __m128i da = _mm_sub_epi16(b, c), db = _mm_sub_epi16(a, c);  // p - a, p - b with p = a + b - c
    const __m128i pa = abs16(da), pb = abs16(db), pc = abs16(_mm_add_epi16(da, db));
    const __m128i smallest = _mm_min_epi16(pc, _mm_min_epi16(pa, pb));
    const __m128i is_a = _mm_cmpeq_epi16(smallest, pa), is_b = _mm_cmpeq_epi16(smallest, pb);  // ties: a, then b, then c
    const __m128i bc = _mm_or_si128(_mm_and_si128(is_b, b), _mm_andnot_si128(is_b, c));
    const __m128i nearest = _mm_or_si128(_mm_and_si128(is_a, a), _mm_andnot_si128(is_a, bc));
    const __m128i d = _mm_and_si128(_mm_add_epi16(x, nearest), low8);
    px_store<BPP>(cur + i, d);
    c = b;
    a = d;
  }
}

template <int BPP>
void average_row_simd(uint8_t* cur, const uint8_t* prev, size_t n) {
  const __m128i low8 = _mm_set1_epi16(0xFF);
  __m128i a = _mm_setzero_si128();
  for (size_t i = 0; i + BPP <= n; i += BPP) {
    const __m128i b = px_load(prev + i), x = px_load(cur + i);
    const __m128i d = _mm_and_si128(_mm_add_epi16(x, _mm_srli_epi16(_mm_add_epi16(a, b), 1)), low8);
    px_store<BPP>(cur + i, d);
    a = d;
  }
}
#endif

// reverses the scan-line filter in place; `prev` is the reconstructed line above (nullptr for the first line)
template <int BPP>
void unfilter_row(int filter, uint8_t* __restrict__ cur, const uint8_t* __restrict__ prev, size_t n) {
  switch (filter) {
    case 1:
      for (size_t i = BPP; i < n; ++i) cur[i] = uint8_t(cur[i] + cur[i - BPP]);
      break;
    case 2:
      if (prev)
        for (size_t i = 0; i < n; ++i) cur[i] = uint8_t(cur[i] + prev[i]);
      break;
    case 3:
#if defined(__SSE2__)
      if (prev && (BPP == 3 || BPP == 4) && n % BPP == 0) {
        average_row_simd<BPP>(cur, prev, n);
        break;
      }
#endif
      if (prev) {
        for (size_t i = 0; i < BPP && i < n; ++i) cur[i] = uint8_t(cur[i] + (prev[i] >> 1));
        for (size_t i = BPP; i < n; ++i) cur[i] = uint8_t(cur[i] + ((cur[i - BPP] + prev[i]) >> 1));
      } else {
        for (size_t i = BPP; i < n; ++i) cur[i] = uint8_t(cur[i] + (cur[i - BPP] >> 1));
      }
      break;
    case 4:
#if defined(__SSE2__)
      if (prev && (BPP == 3 || BPP == 4) && n % BPP == 0) {
        paeth_row_simd<BPP>(cur, prev, n);
        break;
      }
#endif
      if (prev) {
        for (size_t i = 0; i < BPP && i < n; ++i) cur[i] = uint8_t(cur[i] + prev[i]);
        for (size_t i = BPP; i < n; ++i) cur[i] = uint8_t(cur[i] + paeth(cur[i - BPP], prev[i], prev[i - BPP]));
      } else {
        for (size_t i = BPP; i < n; ++i) cur[i] = uint8_t(cur[i] + cur[i - BPP]);
      }
      break;
    default: break;
  }
}

void unfilter_dispatch(int bpp, int filter, uint8_t* cur, const uint8_t* prev, size_t n) {
  switch (bpp) {
    case 1: unfilter_row<1>(filter, cur, prev, n); break;
    case 2: unfilter_row<2>(filter, cur, prev, n); break;
    case 3: unfilter_row<3>(filter, cur, prev, n); break;
    case 4: unfilter_row<4>(filter, cur, prev, n); break;
    case 6: unfilter_row<6>(filter, cur, prev, n); break;
    default: unfilter_row<8>(filter, cur, prev, n); break;
  }
}

// one reconstructed scan line -> W pixels of 3 bytes in the requested channel order
void convert_row(const PngHeader& hd, const uint8_t* row, const uint8_t* plte, int n_plte, int bgr, uint8_t* out, bool* bad_index) {
  const int W = int(hd.w);
  const int r = bgr ? 2 : 0, b = bgr ? 0 : 2;
  const int step = hd.depth == 16 ? 2 : 1;  // 16-bit samples are big-endian: the high byte comes first
  switch (hd.ctype) {
    case 2:
    case 6: {
      const int px = (hd.ctype == 2 ? 3 : 4) * step;
      if (px == 3 && !bgr) {
        memcpy(out, row, size_t(W) * 3);
      } else {
        for (int x = 0; x < W; ++x) {
          const uint8_t* s = row + size_t(x) * px;
          out[3 * x + r] = s[0];
          out[3 * x + 1] = s[step];
          out[3 * x + b] = s[2 * step];
        }
      }
    } break;
    case 4: {
      const int px = 2 * step;
      for (int x = 0; x < W; ++x) {
        const uint8_t v = row[size_t(x) * px];
        out[3 * x] = out[3 * x + 1] = out[3 * x + 2] = v;
      }
    } break;
    case 0:
    case 3: {
      const int dp = hd.depth;
      const int mul = dp == 1 ? 255 : dp == 2 ? 85 : dp == 4 ? 17 : 1;
      for (int x = 0; x < W; ++x) {
        int v;
        if (dp >= 8) {
          v = row[size_t(x) * step];
        } else {
          const int per = 8 / dp, shift = (per - 1 - (x % per)) * dp;
          v = (row[x / per] >> shift) & ((1 << dp) - 1);
        }
        if (hd.ctype == 0) {
          const uint8_t g = uint8_t(dp < 8 ? v * mul : v);
          out[3 * x] = out[3 * x + 1] = out[3 * x + 2] = g;
        } else {
          if (v >= n_plte) {
            *bad_index = true;
            v = 0;
          }
          out[3 * x + r] = plte[3 * v];
          out[3 * x + 1] = plte[3 * v + 1];
          out[3 * x + b] = plte[3 * v + 2];
        }
      }
    } break;
    default: break;
  }
}

int decode_png(const uint8_t* png, size_t len, int H, int W, int order, uint8_t* out, size_t row_pitch, Scratch* sc, std::string* err) {
  PngHeader hd;
  int rc = parse_header(png, len, &hd, err);
  if (rc != BP_OK) return rc;
  if (hd.interlace) {
    *err = "PNG: Adam7-interlaced files are not decoded by the native ingest";
    return BP_ERR_UNSUPPORTED;
  }
  if (int(hd.h) != H || int(hd.w) != W) {
    *err = "PNG: frame is " + std::to_string(hd.w) + "x" + std::to_string(hd.h) + ", expected " + std::to_string(W) + "x" + std::to_string(H);
    return BP_ERR_INVALID;
  }
  if (row_pitch < size_t(W) * 3) {
    *err = "PNG: row pitch smaller than a row";
    return BP_ERR_INVALID;
  }
  const size_t bits = size_t(hd.channels()) * hd.depth;
  const size_t row_bytes = (size_t(W) * bits + 7) / 8;
  const int bpp = int(bits >= 8 ? bits / 8 : 1);
  const size_t raw_bytes = (row_bytes + 1) * H;
  sc->raw.resize(raw_bytes + bp_inflate::kOutPad);
  sc->comp.clear();
  sc->comp.reserve(len + bp_inflate::kInPad);  // the IDAT payloads, concatenated: one zlib stream

  uint8_t plte[768];
  int n_plte = 0;
  bool seen_end = false, seen_idat = false;
  size_t pos = 8;
  while (pos + 12 <= len && !seen_end) {
    const uint32_t clen = be32(png + pos);
    const uint8_t* type = png + pos + 4;
    if (clen > 0x7fffffffu || pos + 12 + size_t(clen) > len) {
      *err = "PNG: truncated chunk";
      return BP_ERR_INVALID;
    }
    const uint8_t* data = type + 4;
    const bool critical = !(type[0] & 0x20);
    if (critical) {  // as libpng: a CRC error in a critical chunk is fatal, ancillary chunks are skipped unchecked
      const uint32_t crc = uint32_t(crc32(crc32(0L, Z_NULL, 0), type, uInt(clen + 4)));
      if (crc != be32(data + clen)) {
        *err = "PNG: CRC mismatch in a critical chunk";
        return BP_ERR_INVALID;
      }
    }
    if (!memcmp(type, "IDAT", 4)) {
      seen_idat = true;
      sc->comp.insert(sc->comp.end(), data, data + clen);
    } else if (!memcmp(type, "PLTE", 4)) {
      if (clen % 3 != 0 || clen > 768 || seen_idat) {
        *err = "PNG: bad PLTE";
        return BP_ERR_INVALID;
      }
      memcpy(plte, data, clen);
      n_plte = int(clen / 3);
    } else if (!memcmp(type, "IEND", 4)) {
      seen_end = true;
    } else if (critical && memcmp(type, "IHDR", 4) != 0) {
      *err = "PNG: unknown critical chunk";
      return BP_ERR_INVALID;
    }
    pos += 12 + size_t(clen);
  }
  const size_t comp_len = sc->comp.size();
  sc->comp.resize(comp_len + bp_inflate::kInPad, 0);
  size_t produced = 0;
  const int zrc = bp_inflate::inflate_zlib(sc->comp.data(), comp_len, sc->raw.data(), raw_bytes, &produced, &sc->tables);
  if (zrc != bp_inflate::OK) {
    *err = std::string("PNG: image data: ") + bp_inflate::status_text(zrc) + (seen_end ? "" : " (file truncated?)");
    return BP_ERR_INVALID;
  }
  if (produced != raw_bytes) {
    *err = "PNG: less image data than the header announces";
    return BP_ERR_INVALID;
  }
  if (hd.ctype == 3 && n_plte == 0) {
    *err = "PNG: palette image without PLTE";
    return BP_ERR_INVALID;
  }

  const int bgr = order == BP_ORDER_BGR;
  bool bad_index = false;
  const uint8_t* prev = nullptr;
  for (int y = 0; y < H; ++y) {
    uint8_t* line = sc->raw.data() + size_t(y) * (row_bytes + 1);
    const int filter = line[0];
    if (filter > 4) {
      *err = "PNG: unknown filter type";
      return BP_ERR_INVALID;
    }
    unfilter_dispatch(bpp, filter, line + 1, prev, row_bytes);
    prev = line + 1;
    convert_row(hd, line + 1, plte, n_plte, bgr, out + size_t(y) * row_pitch, &bad_index);
  }
  if (bad_index) {
    *err = "PNG: palette index out of range";
    return BP_ERR_INVALID;
  }
  return BP_OK;
}

// ---- containers without entropy coding: a decoded-once copy of a sequence that the pool can deliver at memory speed.
// Binary PPM / PGM ("P6" / "P5", maxval 255) and NumPy .npy files holding uint8 [H, W, 3] (C order).
inline void copy_rgb_rows(const uint8_t* src, int H, int W, int channels, int bgr, uint8_t* out, size_t row_pitch) {
  for (int y = 0; y < H; ++y) {
    const uint8_t* s = src + size_t(y) * W * channels;
    uint8_t* d = out + size_t(y) * row_pitch;
    if (channels == 3 && !bgr) {
      memcpy(d, s, size_t(W) * 3);
    } else if (channels == 3) {
      for (int x = 0; x < W; ++x) {
        d[3 * x] = s[3 * x + 2];
        d[3 * x + 1] = s[3 * x + 1];
        d[3 * x + 2] = s[3 * x];
      }
    } else {
      for (int x = 0; x < W; ++x) d[3 * x] = d[3 * x + 1] = d[3 * x + 2] = s[x];
    }
  }
}

int decode_pnm(const uint8_t* p, size_t len, int H, int W, int order, uint8_t* out, size_t row_pitch, std::string* err) {
  const int channels = p[1] == '6' ? 3 : 1;
  size_t pos = 2;
  long vals[3];
  for (int k = 0; k < 3; ++k) {  // width, height, maxval: decimal numbers separated by white space, '#' starts a comment
    for (;;) {
      while (pos < len && (p[pos] == ' ' || p[pos] == '\t' || p[pos] == '\n' || p[pos] == '\r')) ++pos;
      if (pos < len && p[pos] == '#') {
        while (pos < len && p[pos] != '\n') ++pos;
        continue;
      }
      break;
    }
    if (pos >= len || p[pos] < '0' || p[pos] > '9') {
      *err = "PNM: bad header";
      return BP_ERR_INVALID;
    }
    long v = 0;
    while (pos < len && p[pos] >= '0' && p[pos] <= '9' && v < (1L << 24)) v = v * 10 + (p[pos++] - '0');
    vals[k] = v;
  }
  if (pos >= len || !(p[pos] == ' ' || p[pos] == '\t' || p[pos] == '\n' || p[pos] == '\r')) {
    *err = "PNM: bad header";
    return BP_ERR_INVALID;
  }
  ++pos;  // exactly one white-space byte before the samples
  if (vals[2] != 255) {
    *err = "PNM: only maxval 255 is read by the native ingest";
    return BP_ERR_UNSUPPORTED;
  }
  if (vals[0] != W || vals[1] != H) {
    *err = "PNM: frame is " + std::to_string(vals[0]) + "x" + std::to_string(vals[1]) + ", expected " + std::to_string(W) + "x" + std::to_string(H);
    return BP_ERR_INVALID;
  }
  if (len - pos < size_t(H) * W * channels) {
    *err = "PNM: sample data ends early (truncated file?)";
    return BP_ERR_INVALID;
  }
  copy_rgb_rows(p + pos, H, W, channels, order == BP_ORDER_BGR, out, row_pitch);
  return BP_OK;
}

int decode_npy(const uint8_t* p, size_t len, int H, int W, int order, uint8_t* out, size_t row_pitch, std::string* err) {
  if (len < 12) {
    *err = "NPY: truncated header";
    return BP_ERR_INVALID;
  }
  const int major = p[6];
  const size_t hlen = major == 1 ? size_t(p[8] | (p[9] << 8)) : size_t(p[8]) | (size_t(p[9]) << 8) | (size_t(p[10]) << 16) | (size_t(p[11]) << 24);
  const size_t hoff = major == 1 ? 10 : 12;
  if (major < 1 || major > 3 || hoff + hlen > len) {
    *err = "NPY: bad header";
    return BP_ERR_INVALID;
  }
  const std::string head(reinterpret_cast<const char*>(p + hoff), hlen);
  const std::string shape3 = "(" + std::to_string(H) + ", " + std::to_string(W) + ", 3)";
  const std::string shape1 = "(" + std::to_string(H) + ", " + std::to_string(W) + ")";
  const bool u1 = head.find("'|u1'") != std::string::npos || head.find("'<u1'") != std::string::npos || head.find("'u1'") != std::string::npos;
  const bool c_order = head.find("'fortran_order': False") != std::string::npos;
  const bool rgb = head.find("'shape': " + shape3) != std::string::npos, grey = head.find("'shape': " + shape1) != std::string::npos;
  if (!u1 || !c_order) {
    *err = "NPY: only C-ordered uint8 arrays are read by the native ingest";
    return BP_ERR_UNSUPPORTED;
  }
  if (!rgb && !grey) {
    *err = "NPY: array shape is not " + shape3 + " or " + shape1;
    return BP_ERR_INVALID;
  }
  const int channels = rgb ? 3 : 1;
  if (len - (hoff + hlen) < size_t(H) * W * channels) {
    *err = "NPY: array data ends early (truncated file?)";
    return BP_ERR_INVALID;
  }
  copy_rgb_rows(p + hoff + hlen, H, W, channels, order == BP_ORDER_BGR, out, row_pitch);
  return BP_OK;
}

// PNG, binary PPM / PGM or .npy, told apart by their magic bytes; anything else is BP_ERR_UNSUPPORTED
int decode_frame(const uint8_t* p, size_t len, int H, int W, int order, uint8_t* out, size_t row_pitch, Scratch* sc, std::string* err) {
  if (len >= 3 && p[0] == 'P' && (p[1] == '6' || p[1] == '5')) return decode_pnm(p, len, H, W, order, out, row_pitch, err);
  if (len >= 6 && memcmp(p, "\x93NUMPY", 6) == 0) return decode_npy(p, len, H, W, order, out, row_pitch, err);
  return decode_png(p, len, H, W, order, out, row_pitch, sc, err);
}

int read_file(const char* path, std::vector<uint8_t>* buf, std::string* err) {
  const int fd = open(path, O_RDONLY | O_CLOEXEC);
  if (fd < 0) {
    *err = std::string("cannot open ") + path;
    return BP_ERR_IO;
  }
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size <= 0) {
    close(fd);
    *err = std::string("cannot stat / empty file ") + path;
    return BP_ERR_IO;
  }
  buf->resize(size_t(st.st_size));
  size_t got = 0;
  while (got < buf->size()) {
    const ssize_t r = read(fd, buf->data() + got, buf->size() - got);
    if (r <= 0) break;
    got += size_t(r);
  }
  close(fd);
  if (got != buf->size()) {
    *err = std::string("short read on ") + path;
    return BP_ERR_IO;
  }
  return BP_OK;
}

struct Batch {
  int remaining = 0;
  int first_error = BP_OK;
  std::string message;
};

struct Job {
  int64_t ticket;
  std::string path;
  uint8_t* out;
  int H, W, order;
  int32_t* status;
};

}  // namespace

struct bp_ingest {
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  std::deque<Job> jobs;
  std::map<int64_t, Batch> batches;
  int64_t next_ticket = 1;
  bool stop = false;

  void work() {
    Scratch sc;
    for (;;) {
      Job j;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_job.wait(lk, [&] { return stop || !jobs.empty(); });
        if (jobs.empty()) return;  // stop requested and nothing left
        j = std::move(jobs.front());
        jobs.pop_front();
      }
      std::string err;
      int rc = read_file(j.path.c_str(), &sc.file, &err);
      if (rc == BP_OK) rc = decode_frame(sc.file.data(), sc.file.size(), j.H, j.W, j.order, j.out, size_t(j.W) * 3, &sc, &err);
      if (j.status) *j.status = rc;
      {
        std::lock_guard<std::mutex> lk(mu);
        Batch& b = batches[j.ticket];
        if (rc != BP_OK && b.first_error == BP_OK) {
          b.first_error = rc;
          b.message = j.path + ": " + err;
        }
        if (--b.remaining == 0) cv_done.notify_all();
      }
    }
  }
};

extern "C" {

int bp_png_info(const uint8_t* png, size_t len, int* H, int* W, int* channels, int* depth) {
  if (!png) return bp_fail(BP_ERR_INVALID, "bp_png_info: null stream");
  PngHeader hd;
  std::string err;
  const int rc = parse_header(png, len, &hd, &err);
  if (rc != BP_OK) return bp_fail(rc, err.c_str());
  if (H) *H = int(hd.h);
  if (W) *W = int(hd.w);
  if (channels) *channels = hd.ctype == 3 ? 3 : hd.channels();
  if (depth) *depth = hd.depth;
  return BP_OK;
}

int bp_png_decode(const uint8_t* png, size_t len, int H, int W, int order, uint8_t* out, size_t row_pitch) {
  if (!png || !out || H <= 0 || W <= 0 || (order != BP_ORDER_RGB && order != BP_ORDER_BGR))
    return bp_fail(BP_ERR_INVALID, "bp_png_decode: bad arguments");
  static thread_local Scratch sc;
  std::string err;
  const int rc = decode_png(png, len, H, W, order, out, row_pitch ? row_pitch : size_t(W) * 3, &sc, &err);
  return rc == BP_OK ? BP_OK : bp_fail(rc, err.c_str());
}

int bp_frame_decode(const uint8_t* data, size_t len, int H, int W, int order, uint8_t* out, size_t row_pitch) {
  if (!data || !out || H <= 0 || W <= 0 || (order != BP_ORDER_RGB && order != BP_ORDER_BGR))
    return bp_fail(BP_ERR_INVALID, "bp_frame_decode: bad arguments");
  if (row_pitch && row_pitch < size_t(W) * 3) return bp_fail(BP_ERR_INVALID, "bp_frame_decode: row pitch smaller than a row");
  static thread_local Scratch sc;
  std::string err;
  const int rc = decode_frame(data, len, H, W, order, out, row_pitch ? row_pitch : size_t(W) * 3, &sc, &err);
  return rc == BP_OK ? BP_OK : bp_fail(rc, err.c_str());
}

int bp_zlib_inflate(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_cap, size_t* out_len) {
  if (!in || (!out && out_cap) || !out_len) return bp_fail(BP_ERR_INVALID, "bp_zlib_inflate: null argument");
  static thread_local Scratch sc;  // padded copies: the decoder reads / writes a few bytes past both ends by design
  sc.comp.assign(in, in + in_len);
  sc.comp.resize(in_len + bp_inflate::kInPad, 0);
  sc.raw.resize(out_cap + bp_inflate::kOutPad);
  size_t n = 0;
  const int rc = bp_inflate::inflate_zlib(sc.comp.data(), in_len, sc.raw.data(), out_cap, &n, &sc.tables);
  if (rc != bp_inflate::OK) return bp_fail(BP_ERR_INVALID, (std::string("bp_zlib_inflate: ") + bp_inflate::status_text(rc)).c_str());
  memcpy(out, sc.raw.data(), n);
  *out_len = n;
  return BP_OK;
}

int bp_ingest_create(int n_threads, bp_ingest** out) {
  if (!out) return bp_fail(BP_ERR_INVALID, "bp_ingest_create: null out");
  if (n_threads <= 0) n_threads = int(std::thread::hardware_concurrency());
  if (n_threads <= 0) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  bp_ingest* g = new bp_ingest();
  for (int i = 0; i < n_threads; ++i) g->workers.emplace_back([g] { g->work(); });
  *out = g;
  return BP_OK;
}

void bp_ingest_destroy(bp_ingest* g) {
  if (!g) return;
  {
    std::lock_guard<std::mutex> lk(g->mu);
    g->stop = true;
  }
  g->cv_job.notify_all();
  for (auto& t : g->workers) t.join();  // queued files are still decoded: their output buffers stay valid until here
  delete g;
}

int bp_ingest_num_threads(bp_ingest* g) { return g ? int(g->workers.size()) : 0; }

int64_t bp_ingest_submit(bp_ingest* g, const char* const* paths, int n, int H, int W, int order, uint8_t* out, size_t frame_pitch,
                         int32_t* status) {
  if (!g || !paths || n <= 0 || !out || H <= 0 || W <= 0 || (order != BP_ORDER_RGB && order != BP_ORDER_BGR))
    return bp_fail(BP_ERR_INVALID, "bp_ingest_submit: bad arguments");
  if (frame_pitch == 0) frame_pitch = size_t(H) * W * 3;
  if (frame_pitch < size_t(H) * W * 3) return bp_fail(BP_ERR_INVALID, "bp_ingest_submit: frame pitch smaller than a frame");
  for (int i = 0; i < n; ++i)
    if (!paths[i]) return bp_fail(BP_ERR_INVALID, "bp_ingest_submit: null path");
  int64_t ticket;
  {
    std::lock_guard<std::mutex> lk(g->mu);
    ticket = g->next_ticket++;
    g->batches[ticket].remaining = n;
    for (int i = 0; i < n; ++i) g->jobs.push_back(Job{ticket, paths[i], out + size_t(i) * frame_pitch, H, W, order, status ? status + i : nullptr});
  }
  g->cv_job.notify_all();
  return ticket;
}

int bp_ingest_wait(bp_ingest* g, int64_t ticket) {
  if (!g) return bp_fail(BP_ERR_INVALID, "bp_ingest_wait: null handle");
  std::unique_lock<std::mutex> lk(g->mu);
  auto it = g->batches.find(ticket);
  if (it == g->batches.end()) return bp_fail(BP_ERR_INVALID, "bp_ingest_wait: unknown ticket");
  g->cv_done.wait(lk, [&] { return it->second.remaining == 0; });
  const int rc = it->second.first_error;
  const std::string msg = it->second.message;
  g->batches.erase(it);
  lk.unlock();
  return rc == BP_OK ? BP_OK : bp_fail(rc, msg.c_str());
}

}  // extern "C"
