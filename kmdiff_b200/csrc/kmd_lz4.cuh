// Device-side decoder of LZ4 frame streams as kmtricks writes them (km::lz4_stream, reference
// thirdparty/kmtricks/include/kmtricks/io/lz4_stream.hpp:42, 128-140, 179: LZ4F default preferences = 64 KiB LINKED
// blocks, no checksums; every counts/partition_p/<sample>.kmer.lz4 is a 41-byte header followed by one such frame,
// io/kmer_file.hpp:69-101).
//
// Linked blocks make one stream a serial dependency chain (a match may reach back 64 KiB across block borders), so the
// parallelism is ACROSS streams: one warp per stream, and a run directory in count-file mode holds partitions x samples
// of them — thousands.  Inside a stream the warp parses the sequence headers from a shared-memory window of the input
// (refilled with coalesced 16-byte loads) and copies literals and matches 32 bytes per step; the last 2 KiB of the
// output are mirrored in a shared-memory ring, which serves the near matches sorted (k-mer, count) records consist
// of (the previous record is 12-20 bytes back), far matches read the output back from L2.
//
// Written from the public format descriptions (lz4_Frame_format.md, lz4_Block_format.md), like the host decoder in
// kmdiff_b200/host/lz4_frame.hpp, which is the parity reference of tests/test_gpu_lz4.py.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace kmd {
namespace lz4 {

constexpr int kWarps = 8;            // streams per CTA
constexpr uint32_t kWindow = 1024;   // bytes of input held in shared memory per warp
constexpr uint32_t kRing = 2048;     // bytes of the output tail mirrored in shared memory per warp

enum : uint32_t { kOk = 0, kBadMagic = 1, kBadHeader = 2, kTruncated = 3, kBadOffset = 4, kOverflow = 5 };

struct Args
{
  const uint8_t* in;         // the streams' frames, each starting at a 16-byte aligned offset
  const uint64_t* in_off;    // [n] byte offset of stream s in `in`
  const uint64_t* in_size;   // [n] bytes of stream s
  uint8_t* out;              // null: size pass (sequences are parsed, nothing is copied)
  const uint64_t* out_off;   // [n] byte offset of stream s in `out`
  const uint64_t* out_cap;   // [n] bytes stream s may produce (the size pass's result)
  uint64_t* out_size;        // [n] decoded size
  uint32_t* status;          // [n] kOk or the first error
  uint32_t n;
};

struct Reader
{
  const uint8_t* base;   // stream start in global memory (16-byte aligned)
  uint64_t size;
  uint8_t* win;          // shared-memory window
  uint64_t win_pos;      // stream offset of win[0] (multiple of 16), ~0 = empty
  uint32_t lane;

  // the window takes the stream from pos (rounded down to 16) on; bytes past the stream's end read as 0; warp-collective
  __device__ __forceinline__ void refill(uint64_t pos)
  {
    __syncwarp();
    win_pos = pos & ~15ull;
    for (uint32_t o = lane * 16; o < kWindow; o += 32 * 16)
    {
      const uint64_t at = win_pos + o;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (at < size) v = __ldg(reinterpret_cast<const uint4*>(base + at));   // the buffer is padded to 16 bytes
      *reinterpret_cast<uint4*>(win + o) = v;
    }
    __syncwarp();
  }
  // makes [pos, pos + 1) available
  __device__ __forceinline__ void ensure(uint64_t pos)
  {
    if (pos >= win_pos && pos < win_pos + kWindow) return;
    refill(pos);
  }
  // makes [pos, pos + 32) available
  __device__ __forceinline__ void ensure_span(uint64_t pos)
  {
    if (pos >= win_pos && pos + 32 <= win_pos + kWindow) return;
    refill(pos);
  }
  __device__ __forceinline__ uint32_t byte(uint64_t pos)
  {
    ensure(pos);
    return win[pos - win_pos];
  }
  __device__ __forceinline__ uint32_t le32(uint64_t pos)
  {
    return byte(pos) | (byte(pos + 1) << 8) | (byte(pos + 2) << 16) | (byte(pos + 3) << 24);
  }
};

// one warp decodes (or measures) one stream; all control flow is warp-uniform
__global__ void __launch_bounds__(kWarps * 32) frames_kernel(const Args A)
{
  __shared__ __align__(16) uint8_t s_win[kWarps][kWindow];
  __shared__ uint8_t s_ring[kWarps][kRing];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t s = blockIdx.x * kWarps + warp;
  if (s >= A.n) return;
  Reader R{A.in + A.in_off[s], A.in_size[s], s_win[warp], ~0ull, lane};
  uint8_t* ring = s_ring[warp];
  const bool copy = A.out != nullptr;
  uint8_t* out = copy ? A.out + A.out_off[s] : nullptr;
  const uint64_t cap = copy ? A.out_cap[s] : ~0ull;
  uint64_t ip = 0, op = 0;
  uint32_t err = kOk;

  while (err == kOk && ip + 4 <= R.size)
  {
    // ---- frame header ----
    const uint32_t magic = R.le32(ip);
    ip += 4;
    if ((magic & 0xFFFFFFF0u) == 0x184D2A50u)   // skippable frame
    {
      if (ip + 4 > R.size) { err = kTruncated; break; }
      ip += 4ull + R.le32(ip);
      continue;
    }
    if (magic != 0x184D2204u) { err = kBadMagic; break; }
    if (ip + 3 > R.size) { err = kTruncated; break; }
    const uint32_t flg = R.byte(ip), bd = R.byte(ip + 1);
    ip += 2;
    if ((flg >> 6) != 1 || ((bd >> 4) & 7) < 4) { err = kBadHeader; break; }
    const bool block_checksum = flg & 0x10, content_checksum = flg & 0x04;
    if (flg & 0x08) ip += 8;   // content size
    if (flg & 0x01) ip += 4;   // dictionary id
    ip += 1;                   // header checksum
    const uint64_t frame_start = op;   // matches never reach before the frame (no dictionary)

    // ---- blocks ----
    while (err == kOk)
    {
      if (ip + 4 > R.size) { err = kTruncated; break; }
      const uint32_t hdr = R.le32(ip);
      ip += 4;
      if (hdr == 0)   // EndMark
      {
        if (content_checksum) ip += 4;
        break;
      }
      const uint64_t bsize = hdr & 0x7FFFFFFFu;
      const uint64_t bend = ip + bsize;
      if (bend > R.size) { err = kTruncated; break; }
      if (hdr & 0x80000000u)
      {
        // stored block: bsize literal bytes
        if (op + bsize > cap) { err = kOverflow; break; }
        if (copy)
          for (uint64_t done = 0; done < bsize;)
          {
            R.ensure(ip + done);
            const uint32_t chunk = static_cast<uint32_t>(min(bsize - done, R.win_pos + kWindow - (ip + done)));
            for (uint32_t i = lane; i < chunk; i += 32)
            {
              const uint8_t v = R.win[ip + done + i - R.win_pos];
              out[op + done + i] = v;
              ring[(op + done + i) & (kRing - 1)] = v;
            }
            done += chunk;
          }
        op += bsize;
        ip = bend;
      }
      else
      {
        // compressed block: sequences {token, literals, offset, match}
        while (ip < bend)
        {
          // Short sequences (what sorted records compress into: a few literal bytes + a match of 4..18 bytes a record
          // or two back) run in a tight loop on 32-bit window / ring indices: token, literals and offset come from one
          // stretch of the window, the NEXT token is fetched before the current sequence is copied (the parse chain is
          // one shared-memory load per sequence; the copy runs in its shadow), and ONE warp step writes the literals and
          // the match.  Lane i produces output byte i: a literal, or byte (i - lit) mod off of the `off` bytes before
          // the match, which lie in the ring (earlier sequences) or among this sequence's own literals — no lane waits
          // for another.  Everything else (length extensions, the last sequences of a block, far or long matches,
          // errors) leaves the loop for the general path below.
          if (ip + 20 <= bend)
          {
            R.ensure_span(ip);
            __syncwarp();   // ring bytes written by the general path (literals at a block's end, stored blocks) are read below
            uint32_t w = static_cast<uint32_t>(ip - R.win_pos);
            // last window index a short sequence may start at: its 17 header bytes + the next token lie inside the
            // window, and 20 bytes before the block's end
            const uint64_t room = bend - 20 - R.win_pos;
            const uint32_t w_last = static_cast<uint32_t>(min(room, static_cast<uint64_t>(kWindow - 32)));
            const uint64_t behind = op - frame_start, ahead = cap - op;
            uint32_t avail = behind > 0xFFFF0000ull ? 0xFFFF0000u : static_cast<uint32_t>(behind);   // bytes a match may reach back
            uint32_t budget = ahead > 0xFFFF0000ull ? 0xFFFF0000u : static_cast<uint32_t>(ahead);   // bytes the output may still take
            uint32_t o32 = static_cast<uint32_t>(op), done = 0;
            uint8_t* outp = out + op;
            uint32_t token = R.win[w];
            while (true)
            {
              const uint32_t lit = token >> 4, mcode = token & 15u;
              if (lit == 15 || mcode == 15) break;
              const uint8_t* q = R.win + w;
              const uint32_t off = q[1 + lit] | (static_cast<uint32_t>(q[2 + lit]) << 8);
              const uint32_t ml = mcode + 4, total = lit + ml;
              if (off == 0 || off > avail + lit || off + total > kRing || total > budget) break;
              const uint32_t w_next = w + 3 + lit;
              const uint32_t token_next = R.win[w_next];   // inside the window: w <= kWindow - 32
              if (copy)
              {
                if (lane < total)
                {
                  uint32_t v;
                  if (lane < lit) v = q[1 + lane];
                  else
                  {
                    const uint32_t j = lane - lit;
                    const int32_t sp = static_cast<int32_t>(lit) - static_cast<int32_t>(off) + static_cast<int32_t>(off >= ml ? j : j % off);
                    v = sp >= 0 ? q[1 + sp] : ring[(o32 + static_cast<uint32_t>(sp)) & (kRing - 1)];
                  }
                  outp[done + lane] = static_cast<uint8_t>(v);
                  // ring slots [op, op + total) do not meet the slots read, [op + lit - off, op): off + total <= kRing
                  ring[(o32 + lane) & (kRing - 1)] = static_cast<uint8_t>(v);
                }
                __syncwarp();
              }
              o32 += total;
              done += total;
              avail = min(avail + total, 0xFFFF0000u);
              budget -= total;
              w = w_next;
              token = token_next;
              if (w > w_last) break;
            }
            ip = R.win_pos + w;
            op += done;
            if (done) continue;   // progress was made: come back through the loop head (window refill, block end)
          }
          const uint32_t token = R.byte(ip++);
          uint64_t lit = token >> 4;
          if (lit == 15)
          {
            uint32_t b;
            do { if (ip >= bend) { err = kTruncated; break; } b = R.byte(ip++); lit += b; } while (b == 255);
            if (err) break;
          }
          if (ip + lit > bend) { err = kTruncated; break; }
          if (op + lit > cap) { err = kOverflow; break; }
          if (copy)
            for (uint64_t done = 0; done < lit;)
            {
              R.ensure(ip + done);
              const uint32_t chunk = static_cast<uint32_t>(min(lit - done, R.win_pos + kWindow - (ip + done)));
              for (uint32_t i = lane; i < chunk; i += 32)
              {
                const uint8_t v = R.win[ip + done + i - R.win_pos];
                out[op + done + i] = v;
                ring[(op + done + i) & (kRing - 1)] = v;
              }
              done += chunk;
            }
          ip += lit;
          op += lit;
          if (ip >= bend) break;   // the last sequence of a block has no match
          if (ip + 2 > bend) { err = kTruncated; break; }
          const uint32_t off = R.byte(ip) | (R.byte(ip + 1) << 8);
          ip += 2;
          uint64_t ml = token & 15;
          if (ml == 15)
          {
            uint32_t b;
            do { if (ip >= bend) { err = kTruncated; break; } b = R.byte(ip++); ml += b; } while (b == 255);
            if (err) break;
          }
          ml += 4;
          if (off == 0 || off > op - frame_start) { err = kBadOffset; break; }
          if (op + ml > cap) { err = kOverflow; break; }
          if (copy)
          {
            __syncwarp();   // the bytes this match reads were written by other lanes
            // byte i of the match = byte (i mod off) of the `off` bytes before op: no lane depends on another's write
            if (off + ml <= kRing)
            {
              for (uint64_t i = lane; i < ml; i += 32)
              {
                const uint64_t src = op - off + (off >= ml ? i : i % off);
                const uint8_t v = ring[src & (kRing - 1)];
                out[op + i] = v;
                // written after every lane has read: slots [op - kRing, op - kRing + ml) do not meet [op - off, op)
                ring[(op + i) & (kRing - 1)] = v;
              }
            }
            else
            {
              // far or long match: read the output back (L2), then refresh the ring tail
              for (uint64_t i = lane; i < ml; i += 32)
              {
                const uint64_t src = op - off + (off >= ml ? i : i % off);
                out[op + i] = __ldcg(out + src);
              }
              __syncwarp();
              const uint64_t tail = min(ml, static_cast<uint64_t>(kRing));
              for (uint64_t i = lane; i < tail; i += 32)
                ring[(op + ml - tail + i) & (kRing - 1)] = __ldcg(out + op + ml - tail + i);
            }
            __syncwarp();
          }
          op += ml;
        }
        if (err == kOk && ip != bend) err = kTruncated;
      }
      if (block_checksum) ip += 4;
    }
  }
  if (lane == 0)
  {
    A.out_size[s] = op;
    A.status[s] = err;
  }
}

}  // namespace lz4
}  // namespace kmd
