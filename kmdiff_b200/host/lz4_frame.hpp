// LZ4 frame streams as kmtricks writes them (kmtricks/io/lz4_stream.hpp: LZ4F default preferences =
// 64 KiB linked blocks, no checksums; every `.lz4` file of a run directory and every spill file
// `partitions/p{i}_uncorrected` is one such frame after the kmtricks header).
//
// Written from the public LZ4 frame / block format descriptions (lz4_Frame_format.md,
// lz4_Block_format.md); the reference links the vendored liblz4 1.9.3 instead, which is not available
// to this build.  Reader: any conforming frame (linked or independent blocks, stored blocks, optional
// checksums are skipped, concatenated frames).  Writer: a conforming frame of independent 64 KiB blocks,
// each compressed with a greedy single-probe hash matcher (stored as is when that does not shrink it) —
// readable by liblz4, hence by the reference; spill files are small (<= 1e-3 of the rows), so the
// matcher favours brevity over ratio.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace kmd_host {

// xxHash32 (public specification) — only needed for the one-byte frame-header checksum.
inline uint32_t xxh32(const uint8_t* p, size_t len, uint32_t seed)
{
  const uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
  auto rotl = [](uint32_t x, int r) { return (x << r) | (x >> (32 - r)); };
  auto rd32 = [](const uint8_t* q) { uint32_t v; memcpy(&v, q, 4); return v; };
  const uint8_t* end = p + len;
  uint32_t h;
  if (len >= 16)
  {
    uint32_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
    for (; p + 16 <= end; p += 16)
    {
      v1 = rotl(v1 + rd32(p) * P2, 13) * P1;
      v2 = rotl(v2 + rd32(p + 4) * P2, 13) * P1;
      v3 = rotl(v3 + rd32(p + 8) * P2, 13) * P1;
      v4 = rotl(v4 + rd32(p + 12) * P2, 13) * P1;
    }
    h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
  }
  else
    h = seed + P5;
  h += static_cast<uint32_t>(len);
  for (; p + 4 <= end; p += 4) h = rotl(h + rd32(p) * P3, 17) * P4;
  for (; p < end; p++) h = rotl(h + *p * P5, 11) * P1;
  h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
  return h;
}

// Decodes one LZ4 block into dst[0, cap); matches may reach back before dst (linked blocks: the caller
// guarantees `history` valid bytes precede dst).  Returns the decoded size.
// kLz4Slack bytes past src + n must be readable and kLz4Slack bytes past dst + cap writable: short literal runs and
// matches are copied 16 bytes at a time without trimming the last chunk ("wild copy"), which is where the speed of an
// LZ4 decoder comes from — kmtricks rows are short sequences (a k-mer's repeated high bytes, a few literal counts).
constexpr size_t kLz4Slack = 32;

inline size_t lz4_block_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t history)
{
  const uint8_t* ip = src;
  const uint8_t* const iend = src + n;
  uint8_t* op = dst;
  uint8_t* const oend = dst + cap;
  auto copy16 = [](uint8_t* d, const uint8_t* s_) { uint64_t a, b; memcpy(&a, s_, 8); memcpy(&b, s_ + 8, 8); memcpy(d, &a, 8); memcpy(d + 8, &b, 8); };
  auto copy8 = [](uint8_t* d, const uint8_t* s_) { uint64_t a; memcpy(&a, s_, 8); memcpy(d, &a, 8); };
  // Fast zone: while a whole short sequence (token, <= 14 literals, offset; match of <= 18 bytes) can be read and
  // written without looking at the buffer ends, the per-sequence bounds checks go away and both copies have a fixed
  // size: 16 literal bytes, and 18 match bytes in 8 + 8 + 2 (for offsets >= 8 every chunk reads only bytes written by
  // earlier chunks, which is the replication an overlapping LZ4 match means).  Anything else — a length extension, an
  // offset below 8, an offset before the window, the ends of the block — is ONE sequence of the general code below,
  // after which the fast zone is tried again.
  const bool zone = n > 32 && cap > 64;
  const uint8_t* const ifast = zone ? iend - 18 : src;
  uint8_t* const ofast = zone ? oend - 40 : dst;
  const uint8_t* const lowest = dst - history;
  while (ip < iend)
  {
    while (ip < ifast && op < ofast)
    {
      const unsigned token = *ip;
      const size_t lit = token >> 4, mcode = token & 15;
      if (lit == 15 || mcode == 15) break;
      const uint8_t* const lp = ip + 1;
      const size_t off = lp[lit] | (static_cast<size_t>(lp[lit + 1]) << 8);
      uint8_t* const mp = op + lit;
      if (off < 8 || static_cast<size_t>(mp - lowest) < off) break;
      copy16(op, lp);
      const uint8_t* m = mp - off;
      copy8(mp, m);
      copy8(mp + 8, m + 8);
      uint16_t t2; memcpy(&t2, m + 16, 2); memcpy(mp + 16, &t2, 2);
      ip = lp + lit + 2;
      op = mp + mcode + 4;
    }
    const unsigned token = *ip++;
    size_t lit = token >> 4;
    if (lit == 15)
    {
      unsigned b;
      do { if (ip >= iend) throw std::runtime_error("lz4: truncated literal length"); b = *ip++; lit += b; } while (b == 255);
    }
    if (lit > static_cast<size_t>(iend - ip) || lit > static_cast<size_t>(oend - op)) throw std::runtime_error("lz4: literal run overflows");
    if (lit <= 32)
    {
      copy16(op, ip);                       // within the slack of both buffers
      if (lit > 16) copy16(op + 16, ip + 16);
    }
    else
      memcpy(op, ip, lit);
    op += lit; ip += lit;
    if (ip >= iend) break;  // the last sequence has no match part
    if (iend - ip < 2) throw std::runtime_error("lz4: truncated offset");
    const size_t off = ip[0] | (static_cast<size_t>(ip[1]) << 8);
    ip += 2;
    size_t ml = token & 15;
    if (ml == 15)
    {
      unsigned b;
      do { if (ip >= iend) throw std::runtime_error("lz4: truncated match length"); b = *ip++; ml += b; } while (b == 255);
    }
    ml += 4;
    if (off == 0 || off > static_cast<size_t>(op - dst) + history) throw std::runtime_error("lz4: offset out of window");
    if (ml > static_cast<size_t>(oend - op)) throw std::runtime_error("lz4: match overflows");
    const uint8_t* m = op - off;
    if (off >= 16)
    {
      // source and destination chunks never overlap: 16 bytes at a time, the last chunk may run into the slack
      uint8_t* const end = op + ml;
      do { copy16(op, m); op += 16; m += 16; } while (op < end);
      op = end;
    }
    else if (off >= 8)
    {
      uint8_t* const end = op + ml;
      do { uint64_t a; memcpy(&a, m, 8); memcpy(op, &a, 8); op += 8; m += 8; } while (op < end);
      op = end;
    }
    else for (size_t i = 0; i < ml; i++) *op++ = m[i];  // short-period run
  }
  return static_cast<size_t>(op - dst);
}

// Pull-style reader of a FILE* positioned at the start of an LZ4 frame.
class Lz4FrameReader
{
 public:
  explicit Lz4FrameReader(FILE* f) : m_f(f) {}

  // Up to n decoded bytes into dst; 0 at end of stream.  A block is decoded straight into dst when dst has room for a
  // whole one and already holds the 64 KiB a linked block may refer back to (no copy besides the decoder's own);
  // otherwise — the first blocks of a call, the last one, small reads — through the window, which keeps the history.
  size_t read(uint8_t* dst, size_t n)
  {
    size_t got = 0;
    bool hist_in_dst = false;   // the last blocks went straight to dst: the window's history is stale
    while (got < n)
    {
      if (m_pos < m_end)
      {
        const size_t take = std::min(n - got, m_end - m_pos);
        memcpy(dst + got, m_win.data() + m_pos, take);
        m_pos += take; got += take;
        continue;
      }
      bool stored;
      size_t bs;
      if (!next_header(stored, bs)) break;
      const bool direct = n - got >= m_block_max + kLz4Slack && (!m_linked || got >= kHistory);
      uint8_t* const to = direct ? dst + got : m_win.data() + kHistory;
      if (!direct)
      {
        if (hist_in_dst) { history_from(dst, got); hist_in_dst = false; }
        else
        {
          // slide the last 64 KiB of output in front of the window
          const size_t produced = m_end - kHistory;
          if (m_linked && produced)
          {
            const size_t keep = std::min(kHistory, m_hist + produced);
            memmove(m_win.data() + kHistory - keep, m_win.data() + m_end - keep, keep);
            m_hist = keep;
          }
        }
      }
      size_t out;
      if (stored)
      {
        if (!fill(to, bs)) throw std::runtime_error("lz4: truncated block");
        out = bs;
      }
      else
      {
        if (m_src.size() < bs + kLz4Slack) m_src.resize(bs + kLz4Slack);
        if (!fill(m_src.data(), bs)) throw std::runtime_error("lz4: truncated block");
        out = lz4_block_decode(m_src.data(), bs, to, m_block_max, m_linked ? (direct ? got : m_hist) : 0);
      }
      if (m_block_checksum) { uint32_t c; fill(&c, 4); }
      if (direct) { got += out; hist_in_dst = true; }
      else { m_pos = kHistory; m_end = kHistory + out; }
    }
    if (hist_in_dst) history_from(dst, got);   // the next call starts through the window
    return got;
  }

 private:
  static constexpr size_t kHistory = 64 * 1024;

  bool fill(void* p, size_t n) { return fread(p, 1, n, m_f) == n; }

  // the window's history := the last 64 KiB of the `got` bytes at dst; the window itself holds nothing
  void history_from(const uint8_t* dst, size_t got)
  {
    const size_t keep = std::min(kHistory, got);
    memcpy(m_win.data() + kHistory - keep, dst + got - keep, keep);
    m_hist = keep;
    m_pos = m_end = kHistory;
  }

  bool open_frame()
  {
    uint32_t magic;
    for (;;)
    {
      if (!fill(&magic, 4)) return false;
      if ((magic & 0xFFFFFFF0u) == 0x184D2A50u)  // skippable frame
      {
        uint32_t sz; if (!fill(&sz, 4)) return false;
        if (fseek(m_f, sz, SEEK_CUR)) return false;
        continue;
      }
      break;
    }
    if (magic != 0x184D2204u) throw std::runtime_error("lz4: bad frame magic");
    uint8_t flg, bd;
    if (!fill(&flg, 1) || !fill(&bd, 1)) throw std::runtime_error("lz4: truncated frame header");
    if ((flg >> 6) != 1) throw std::runtime_error("lz4: unsupported frame version");
    m_block_checksum = flg & 0x10;
    m_content_checksum = flg & 0x04;
    m_linked = !(flg & 0x20);
    if (flg & 0x08) { uint64_t cs; fill(&cs, 8); }
    if (flg & 0x01) { uint32_t id; fill(&id, 4); }
    uint8_t hc; fill(&hc, 1);
    const unsigned bs = (bd >> 4) & 7;
    if (bs < 4) throw std::runtime_error("lz4: bad block size code");
    m_block_max = static_cast<size_t>(1) << (8 + 2 * bs);
    m_win.resize(kHistory + m_block_max + kLz4Slack);
    m_hist = 0; m_pos = m_end = kHistory;
    m_in_frame = true;
    return true;
  }

  // the next data block's header (opening the next frame when the current one has ended); false at end of stream
  bool next_header(bool& stored, size_t& n)
  {
    for (;;)
    {
      if (!m_in_frame && !open_frame()) return false;
      uint32_t hdr;
      if (!fill(&hdr, 4)) throw std::runtime_error("lz4: truncated block header");
      if (hdr == 0)  // EndMark
      {
        if (m_content_checksum) { uint32_t c; fill(&c, 4); }
        m_in_frame = false;
        continue;
      }
      stored = hdr & 0x80000000u;
      n = hdr & 0x7FFFFFFFu;
      if (n > m_block_max) throw std::runtime_error("lz4: block larger than the frame's maximum");
      if (n == 0) continue;   // an empty stored block
      return true;
    }
  }

  FILE* m_f;
  std::vector<uint8_t> m_win, m_src;
  size_t m_block_max = 0, m_hist = 0, m_pos = 0, m_end = 0;
  bool m_in_frame = false, m_linked = true, m_block_checksum = false, m_content_checksum = false;
};

// Push-style writer: one frame of stored 64 KiB blocks.
// One LZ4 block (lz4_Block_format.md): sequences of {token, literals, 2-byte offset, match length}; the last five
// bytes are literals and no match starts within the last twelve.  Returns the compressed size (appended to out).
inline size_t lz4_compress_block(const uint8_t* src, size_t n, std::vector<uint8_t>& out)
{
  const size_t start = out.size();
  auto rd32 = [&](size_t i) { uint32_t v; memcpy(&v, src + i, 4); return v; };
  auto put_len = [&](size_t len) { for (; len >= 255; len -= 255) out.push_back(255); out.push_back(static_cast<uint8_t>(len)); };
  auto emit = [&](size_t anchor, size_t lit_end, size_t offset, size_t match_len) {
    const size_t lit = lit_end - anchor;
    const size_t ml = match_len ? match_len - 4 : 0;
    out.push_back(static_cast<uint8_t>((lit >= 15 ? 15 : lit) << 4 | (match_len ? (ml >= 15 ? 15 : ml) : 0)));
    if (lit >= 15) put_len(lit - 15);
    out.insert(out.end(), src + anchor, src + lit_end);
    if (!match_len) return;
    out.push_back(static_cast<uint8_t>(offset & 0xFF));
    out.push_back(static_cast<uint8_t>(offset >> 8));
    if (ml >= 15) put_len(ml - 15);
  };
  size_t anchor = 0;
  if (n >= 13)
  {
    constexpr int kHashBits = 13;
    uint32_t table[1 << kHashBits];
    memset(table, 0xFF, sizeof table);
    const size_t mflimit = n - 12, matchlimit = n - 5;
    for (size_t ip = 0; ip < mflimit;)
    {
      const uint32_t seq = rd32(ip);
      const uint32_t h = (seq * 2654435761u) >> (32 - kHashBits);
      const uint32_t ref = table[h];
      table[h] = static_cast<uint32_t>(ip);
      if (ref != 0xFFFFFFFFu && ip - ref <= 65535 && rd32(ref) == seq)
      {
        size_t ml = 4;
        while (ip + ml < matchlimit && src[ref + ml] == src[ip + ml]) ml++;
        emit(anchor, ip, ip - ref, ml);
        ip += ml;
        anchor = ip;
      }
      else
        ip++;
    }
  }
  emit(anchor, n, 0, 0);
  return out.size() - start;
}

class Lz4FrameWriter
{
 public:
  explicit Lz4FrameWriter(FILE* f) : m_f(f)
  {
    const uint8_t desc[2] = {0x60 /* version 1, independent blocks */, 0x40 /* 64 KiB */};
    const uint32_t magic = 0x184D2204u;
    fwrite(&magic, 4, 1, m_f);
    fwrite(desc, 1, 2, m_f);
    const uint8_t hc = static_cast<uint8_t>(xxh32(desc, 2, 0) >> 8);
    fwrite(&hc, 1, 1, m_f);
    m_buf.reserve(kBlock);
  }
  void write(const void* p, size_t n)
  {
    const uint8_t* s = static_cast<const uint8_t*>(p);
    while (n)
    {
      const size_t take = std::min(n, kBlock - m_buf.size());
      m_buf.insert(m_buf.end(), s, s + take);
      s += take; n -= take;
      if (m_buf.size() == kBlock) flush();
    }
  }
  void close()
  {
    if (!m_f) return;
    flush();
    const uint32_t end = 0;
    fwrite(&end, 4, 1, m_f);
    m_f = nullptr;
  }
  ~Lz4FrameWriter() { close(); }

 private:
  static constexpr size_t kBlock = 64 * 1024;
  void flush()
  {
    if (m_buf.empty()) return;
    m_cmp.clear();
    const size_t c = lz4_compress_block(m_buf.data(), m_buf.size(), m_cmp);
    if (c < m_buf.size())
    {
      const uint32_t hdr = static_cast<uint32_t>(c);
      fwrite(&hdr, 4, 1, m_f);
      fwrite(m_cmp.data(), 1, c, m_f);
    }
    else
    {
      const uint32_t hdr = static_cast<uint32_t>(m_buf.size()) | 0x80000000u;   // stored block
      fwrite(&hdr, 4, 1, m_f);
      fwrite(m_buf.data(), 1, m_buf.size(), m_f);
    }
    m_buf.clear();
  }
  FILE* m_f;
  std::vector<uint8_t> m_buf, m_cmp;
};

}  // namespace kmd_host
