// One-shot DEFLATE / zlib-stream decoder for the frame ingest (csrc/ingest.cu).  Host code.
//
// Why not zlib's inflate(): a 640x480 frame is ~0.5 MB of compressed data and zlib's state-machine decoder spends
// ~7.5 ms on it, 90 % of the whole PNG decode -- the ingest threads, not the GPU, would bound frames-from-disk throughput.
// Knowing the whole input and the exact output size up front (PNG gives both) allows the simpler, faster shape used here:
//   * 64-bit bit buffer refilled with one unaligned 8-byte load, no per-byte input checks (the caller pads the input with
//     kInPad zero bytes and the output with kOutPad bytes of slack; overruns are detected after the fact, never written
//     outside the slack);
//   * 11-bit primary literal/length table + sub-tables, 8-bit primary distance table, entries carrying base value and
//     extra-bit count so a symbol costs one lookup;
//   * up to five literals per refill; matches copied a word at a time (with the overlapping small-distance cases folded
//     into the same loop).
// RFC 1950 / RFC 1951 are the specification; tests/test_ingest.py checks it against Python's zlib on every block type.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace bp_inflate {

constexpr size_t kInPad = 32;    // zero bytes the caller must provide after the last input byte
constexpr size_t kOutPad = 16;   // writable slack the caller must provide after out_cap bytes

enum Status { OK = 0, BAD_HEADER = -1, BAD_BLOCK = -2, BAD_CODE = -3, BAD_DISTANCE = -4, INPUT_ENDS_EARLY = -5, OUTPUT_OVERFLOW = -6,
              BAD_CHECKSUM = -7 };

inline const char* status_text(int s) {
  switch (s) {
    case OK: return "ok";
    case BAD_HEADER: return "bad zlib header";
    case BAD_BLOCK: return "bad block header";
    case BAD_CODE: return "invalid Huffman code";
    case BAD_DISTANCE: return "match distance beyond the start of the output";
    case INPUT_ENDS_EARLY: return "compressed data ends early";
    case OUTPUT_OVERFLOW: return "more data than the output holds";
    case BAD_CHECKSUM: return "adler32 mismatch";
    default: return "error";
  }
}

// table entry layout (uint32):  [31:16] payload  [15] exceptional  [14] literal  [13] sub-table link  [12] end of block
//                               [11:8] extra-bit count (or sub-table index bits)  [7:0] code bits to consume
constexpr uint32_t F_EXC = 1u << 15, F_LIT = 1u << 14, F_SUB = 1u << 13, F_EOB = 1u << 12;
constexpr int kLitBits = 11, kDistBits = 8, kMaxLen = 15;
constexpr int kLitTableSize = (1 << kLitBits) + 288 * 16, kDistTableSize = (1 << kDistBits) + 32 * 128;

#ifdef BP_INFLATE_STATS
static long g_blocks = 0;
#endif
struct Tables {
  uint32_t lit[kLitTableSize];
  uint32_t dist[kDistTableSize];
};

struct ByteReverse {
  uint8_t t[256];
  ByteReverse() {
    for (int v = 0; v < 256; ++v) {
      int r = 0;
      for (int i = 0; i < 8; ++i) r |= ((v >> i) & 1) << (7 - i);
      t[v] = uint8_t(r);
    }
  }
};
// the n low bits of v (n <= 16) in reverse order: Huffman codes are packed most-significant bit first
inline uint32_t reverse_bits(uint32_t v, int n) {
  static const ByteReverse br;
  return ((uint32_t(br.t[v & 0xFF]) << 8) | br.t[(v >> 8) & 0xFF]) >> (16 - n);
}

// Canonical-code decode table.  lens[n]: code length per symbol (0 = unused); payload[sym] = entry without its bit count.
// Unassigned slots decode as an error.  Returns false for an over-subscribed code.
inline bool build_table(const uint8_t* lens, int n, const uint32_t* payload, int primary_bits, uint32_t* table, int table_cap) {
  int count[kMaxLen + 1] = {0};
  for (int i = 0; i < n; ++i) count[lens[i]]++;
  count[0] = 0;
  uint32_t next_code[kMaxLen + 2];
  uint32_t code = 0;
  int64_t left = 1;
  for (int l = 1; l <= kMaxLen; ++l) {
    left = (left << 1) - count[l];
    if (left < 0) return false;
    code = (code + uint32_t(count[l - 1])) << 1;
    next_code[l] = code;
  }
  const int primary = 1 << primary_bits;
  for (int i = 0; i < primary; ++i) table[i] = F_EXC;  // exceptional without SUB / EOB = invalid code
  // pass 1: longest code behind every primary slot that needs a sub-table
  uint8_t sub_bits[1 << kLitBits];
  memset(sub_bits, 0, size_t(primary));
  uint32_t nc[kMaxLen + 2];
  memcpy(nc, next_code, sizeof(nc));
  for (int s = 0; s < n; ++s) {
    const int l = lens[s];
    if (l <= primary_bits) {
      if (l) nc[l]++;
      continue;
    }
    const uint32_t rev = reverse_bits(nc[l]++, l);
    uint8_t& sb = sub_bits[rev & uint32_t(primary - 1)];
    if (l - primary_bits > sb) sb = uint8_t(l - primary_bits);
  }
  int used = primary;
  for (int i = 0; i < primary; ++i) {
    if (!sub_bits[i]) continue;
    const int size = 1 << sub_bits[i];
    if (used + size > table_cap) return false;
    table[i] = (uint32_t(used) << 16) | F_EXC | F_SUB | (uint32_t(sub_bits[i]) << 8) | uint32_t(primary_bits);
    for (int k = 0; k < size; ++k) table[used + k] = F_EXC;
    used += size;
  }
  // pass 2: fill
  for (int s = 0; s < n; ++s) {
    const int l = lens[s];
    if (!l) continue;
    const uint32_t rev = reverse_bits(next_code[l]++, l);
    if (l <= primary_bits) {
      const uint32_t e = payload[s] | uint32_t(l);
      for (uint32_t k = rev; k < uint32_t(primary); k += 1u << l) table[k] = e;
    } else {
      const uint32_t link = table[rev & uint32_t(primary - 1)];
      const int sb = int((link >> 8) & 0xF), base = int(link >> 16), sl = l - primary_bits;
      const uint32_t e = payload[s] | uint32_t(sl);
      for (uint32_t k = rev >> primary_bits; k < (1u << sb); k += 1u << sl) table[base + k] = e;
    }
  }
  return true;
}

struct Payloads {
  uint32_t lit[288], dist[32], pre[19];
  Payloads() {
    static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    for (int s = 0; s < 256; ++s) lit[s] = (uint32_t(s) << 16) | F_LIT;
    lit[256] = F_EXC | F_EOB;
    for (int s = 257; s < 286; ++s) lit[s] = (uint32_t(lbase[s - 257]) << 16) | (uint32_t(lext[s - 257]) << 8);
    lit[286] = lit[287] = F_EXC;  // codes that take part in the fixed code but never appear in valid data
    for (int s = 0; s < 30; ++s) dist[s] = (uint32_t(dbase[s]) << 16) | (uint32_t(dext[s]) << 8);
    dist[30] = dist[31] = F_EXC;
    for (int s = 0; s < 19; ++s) pre[s] = uint32_t(s) << 16;
  }
};

inline uint64_t load64(const uint8_t* p) {
  uint64_t v;
  memcpy(&v, p, 8);
  return v;  // little-endian hosts only (x86-64 / aarch64), as the rest of this library
}
inline void store64(uint8_t* p, uint64_t v) { memcpy(p, &v, 8); }

// Raw DEFLATE (RFC 1951).  in[0..in_len) followed by kInPad readable zero bytes; out[0..out_cap) followed by kOutPad
// writable bytes.  *out_len = bytes produced, *in_used = whole bytes consumed.
inline int inflate_raw(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_cap, size_t* out_len, size_t* in_used, Tables* tb) {
  static const Payloads pl;
  const uint8_t* ip = in;
  const uint8_t* const in_end = in + in_len;
  uint8_t* op = out;
  uint8_t* const out_end = out + out_cap;
  uint64_t bits = 0;
  unsigned cnt = 0;
#define BP_REFILL()                                   \
  do {                                                \
    bits |= load64(ip) << cnt;                        \
    ip += (63 - cnt) >> 3;                            \
    cnt |= 56;                                        \
  } while (0)
#define BP_TAKE(n) (bits >>= ((n) & 63), cnt -= (n))  // & 63: lets the compiler use the register shift as is
// first input byte no consumed bit belongs to.  The bit buffer looks ahead of it by up to 8 bytes, and one round of the
// symbol loop consumes at most 13 more, so reads stay within kInPad of in_end as long as this is checked once per round.
#define BP_CONSUMED() (ip - (cnt >> 3))
  bool last = false;
  while (!last) {
    if (BP_CONSUMED() > in_end) return INPUT_ENDS_EARLY;
    BP_REFILL();
    last = bits & 1;
    const unsigned type = unsigned(bits >> 1) & 3;
    BP_TAKE(3);
    if (type == 0) {  // stored
      BP_TAKE(cnt & 7);
      ip -= cnt >> 3;  // give whole unread bytes back
      bits = 0;
      cnt = 0;
      if (in_end - ip < 4) return INPUT_ENDS_EARLY;
      const unsigned len = ip[0] | (ip[1] << 8), nlen = ip[2] | (ip[3] << 8);
      if ((len ^ 0xFFFFu) != nlen) return BAD_BLOCK;
      ip += 4;
      if (size_t(in_end - ip) < len) return INPUT_ENDS_EARLY;
      if (size_t(out_end - op) < len) return OUTPUT_OVERFLOW;
      memcpy(op, ip, len);
      ip += len;
      op += len;
      continue;
    }
    if (type == 3) return BAD_BLOCK;
    uint8_t lens[288 + 32];
    int nlit, ndist;
    if (type == 1) {
      nlit = 288;
      ndist = 32;
      for (int i = 0; i < 144; ++i) lens[i] = 8;
      for (int i = 144; i < 256; ++i) lens[i] = 9;
      for (int i = 256; i < 280; ++i) lens[i] = 7;
      for (int i = 280; i < 288; ++i) lens[i] = 8;
      for (int i = 0; i < 32; ++i) lens[288 + i] = 5;
    } else {
      nlit = int(bits & 31) + 257;
      ndist = int((bits >> 5) & 31) + 1;
      const int npre = int((bits >> 10) & 15) + 4;
      BP_TAKE(14);
      if (nlit > 286 || ndist > 30) return BAD_BLOCK;
      static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
      uint8_t plen[19] = {0};
      for (int i = 0; i < npre; ++i) {
        if (cnt < 3) BP_REFILL();
        plen[order[i]] = uint8_t(bits & 7);
        BP_TAKE(3);
      }
      uint32_t ptab[128];
      if (!build_table(plen, 19, pl.pre, 7, ptab, 128)) return BAD_CODE;
      int i = 0;
      const int total = nlit + ndist;
      while (i < total) {
        if (BP_CONSUMED() > in_end) return INPUT_ENDS_EARLY;
        BP_REFILL();
        const uint32_t e = ptab[bits & 127];
        if (e & F_EXC) return BAD_CODE;
        BP_TAKE(e & 0xFF);
        const int sym = int(e >> 16);
        if (sym < 16) {
          lens[i++] = uint8_t(sym);
          continue;
        }
        int rep;
        uint8_t v = 0;
        if (sym == 16) {
          if (i == 0) return BAD_CODE;
          v = lens[i - 1];
          rep = 3 + int(bits & 3);
          BP_TAKE(2);
        } else if (sym == 17) {
          rep = 3 + int(bits & 7);
          BP_TAKE(3);
        } else {
          rep = 11 + int(bits & 127);
          BP_TAKE(7);
        }
        if (i + rep > total) return BAD_CODE;
        memset(lens + i, v, size_t(rep));
        i += rep;
      }
      if (lens[256] == 0) return BAD_CODE;  // no end-of-block code
      memmove(lens + 288, lens + nlit, size_t(ndist));  // distance lengths to their fixed place
      memset(lens + nlit, 0, size_t(288 - nlit));
      memset(lens + 288 + ndist, 0, size_t(32 - ndist));
      nlit = 288;
      ndist = 32;
    }
#ifdef BP_INFLATE_STATS
    ++g_blocks;
#endif
    if (!build_table(lens, nlit, pl.lit, kLitBits, tb->lit, kLitTableSize)) return BAD_CODE;
    if (!build_table(lens + 288, ndist, pl.dist, kDistBits, tb->dist, kDistTableSize)) return BAD_CODE;
    const uint32_t* const lt = tb->lit;
    const uint32_t* const dt = tb->dist;

    for (;;) {
      if (BP_CONSUMED() > in_end) return INPUT_ENDS_EARLY;
      if (op > out_end) return OUTPUT_OVERFLOW;
      BP_REFILL();
      uint32_t e = lt[bits & ((1u << kLitBits) - 1)];
      // up to five literals on one refill: a literal found in the primary table is at most kLitBits long, 5 x 11 <= 56
#define BP_LITERAL()                 \
  BP_TAKE(e & 0xFF);                 \
  *op++ = uint8_t(e >> 16);          \
  e = lt[bits & ((1u << kLitBits) - 1)]
      if (e & F_LIT) {
        BP_LITERAL();
        if (e & F_LIT) {
          BP_LITERAL();
          if (e & F_LIT) {
            BP_LITERAL();
            if (e & F_LIT) {
              BP_LITERAL();
              if (e & F_LIT) {
                BP_TAKE(e & 0xFF);
                *op++ = uint8_t(e >> 16);
                continue;
              }
            }
          }
        }
        BP_REFILL();
      }
#undef BP_LITERAL
      if (e & F_EXC) {
        if (e & F_SUB) {
          BP_TAKE(e & 0xFF);
          e = lt[(e >> 16) + (bits & ((1u << ((e >> 8) & 0xF)) - 1))];
          if (e & F_LIT) {
            BP_TAKE(e & 0xFF);
            *op++ = uint8_t(e >> 16);
            continue;
          }
        }
        if (e & F_EXC) {
          if (e & F_EOB) {
            BP_TAKE(e & 0xFF);
            break;
          }
          return BAD_CODE;
        }
      }
      // length symbol: <= 15 + 5 bits, then distance: <= 15 + 13 bits; 48 <= 56 available after the refill above
      BP_TAKE(e & 0xFF);
      const unsigned lx = (e >> 8) & 0xF;
      const size_t len = (e >> 16) + (bits & ((1u << lx) - 1));
      BP_TAKE(lx);
      uint32_t d = dt[bits & ((1u << kDistBits) - 1)];
      if (d & F_EXC) {
        if (!(d & F_SUB)) return BAD_CODE;
        BP_TAKE(d & 0xFF);
        d = dt[(d >> 16) + (bits & ((1u << ((d >> 8) & 0xF)) - 1))];
        if (d & F_EXC) return BAD_CODE;
      }
      BP_TAKE(d & 0xFF);
      const unsigned dx = (d >> 8) & 0xF;
      const size_t dist = (d >> 16) + (bits & ((1u << dx) - 1));
      BP_TAKE(dx);
      if (dist > size_t(op - out)) return BAD_DISTANCE;
      if (ptrdiff_t(len) > out_end - op) return OUTPUT_OVERFLOW;  // signed: op may be 1-2 literals past out_end here
      const uint8_t* src = op - dist;
      uint8_t* const end = op + len;
      if (dist >= 8) {
        do {
          store64(op, load64(src));
          op += 8;
          src += 8;
        } while (op < end);
      } else if (dist == 1) {
        const uint64_t v = 0x0101010101010101ull * src[0];
        do {
          store64(op, v);
          op += 8;
        } while (op < end);
      } else {
        // overlapping copy: each store lays down `dist` final bytes; the tail of the word is rewritten by the next one
        do {
          store64(op, load64(src));
          op += dist;
          src += dist;
        } while (op < end);
      }
      op = end;
    }
  }
  if (op > out_end) return OUTPUT_OVERFLOW;
  const uint8_t* consumed = BP_CONSUMED();  // whole bytes still sitting unread in the bit buffer are given back
#undef BP_REFILL
#undef BP_TAKE
#undef BP_CONSUMED
  if (consumed > in_end) return INPUT_ENDS_EARLY;
  *out_len = size_t(op - out);
  *in_used = size_t(consumed - in);
  return OK;
}

inline uint32_t adler32(const uint8_t* p, size_t n) {
  uint32_t a = 1, b = 0;
#if defined(__SSE2__)
  // 16 bytes per step: a += sum(p[i]),  b += 16 * a_before + sum((16 - i) * p[i]); sums kept in vector lanes per block
  const __m128i zero = _mm_setzero_si128();
  const __m128i w_lo = _mm_set_epi16(9, 10, 11, 12, 13, 14, 15, 16), w_hi = _mm_set_epi16(1, 2, 3, 4, 5, 6, 7, 8);
  while (n >= 16) {
    size_t chunks = n / 16;
    if (chunks > 346) chunks = 346;  // 5536 bytes: nothing below can overflow before the modulo
    n -= chunks * 16;
    __m128i s1 = zero, prev = zero, s2 = zero;  // s1, prev: 2 x u64; s2: 4 x u32
    for (size_t c = 0; c < chunks; ++c, p += 16) {
      const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(p));
      prev = _mm_add_epi64(prev, s1);
      s1 = _mm_add_epi64(s1, _mm_sad_epu8(v, zero));
      s2 = _mm_add_epi32(s2, _mm_add_epi32(_mm_madd_epi16(_mm_unpacklo_epi8(v, zero), w_lo), _mm_madd_epi16(_mm_unpackhi_epi8(v, zero), w_hi)));
    }
    uint64_t t1[2], tp[2];
    uint32_t t2[4];
    _mm_storeu_si128(reinterpret_cast<__m128i*>(t1), s1);
    _mm_storeu_si128(reinterpret_cast<__m128i*>(tp), prev);
    _mm_storeu_si128(reinterpret_cast<__m128i*>(t2), s2);
    const uint64_t bb = uint64_t(b) + 16ull * chunks * a + 16ull * (tp[0] + tp[1]) + t2[0] + t2[1] + t2[2] + t2[3];
    a = uint32_t((a + t1[0] + t1[1]) % 65521u);
    b = uint32_t(bb % 65521u);
  }
#endif
  while (n) {
    size_t k = n < 5552 ? n : 5552;  // largest run that cannot overflow 32 bits before the modulo
    n -= k;
    while (k--) {
      a += *p++;
      b += a;
    }
    a %= 65521u;
    b %= 65521u;
  }
  return (b << 16) | a;
}

// zlib stream (RFC 1950): 2-byte header, raw DEFLATE, big-endian adler32 of the output.
inline int inflate_zlib(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_cap, size_t* out_len, Tables* tb, bool verify = true) {
  if (in_len < 6) return INPUT_ENDS_EARLY;
  const unsigned cmf = in[0], flg = in[1];
  if ((cmf & 0x0F) != 8 || (cmf >> 4) > 7 || ((cmf << 8) | flg) % 31 != 0 || (flg & 0x20)) return BAD_HEADER;
  size_t used = 0;
  const int rc = inflate_raw(in + 2, in_len - 2, out, out_cap, out_len, &used, tb);
  if (rc != OK) return rc;
  if (in_len - 2 - used < 4) return INPUT_ENDS_EARLY;
  if (verify) {
    const uint8_t* t = in + 2 + used;
    const uint32_t want = (uint32_t(t[0]) << 24) | (uint32_t(t[1]) << 16) | (uint32_t(t[2]) << 8) | t[3];
    if (adler32(out, *out_len) != want) return BAD_CHECKSUM;
  }
  return OK;
}

}  // namespace bp_inflate
