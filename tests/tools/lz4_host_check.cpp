// Test harness of kmdiff_b200/host/lz4_frame.hpp (no GPU needed):
//   lz4_host_check decode SKIP CHUNK IN OUT   frame after SKIP header bytes of IN, read in pieces of CHUNK bytes -> OUT
//   lz4_host_check encode IN OUT              IN as one frame of this repo's writer -> OUT
#include "lz4_frame.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

int main(int argc, char** argv)
{
  if (argc == 6 && !strcmp(argv[1], "decode"))
  {
    const long skip = atol(argv[2]);
    const size_t chunk = static_cast<size_t>(atol(argv[3]));
    FILE* in = fopen(argv[4], "rb");
    FILE* out = fopen(argv[5], "wb");
    if (!in || !out) return 2;
    fseek(in, skip, SEEK_SET);
    std::vector<uint8_t> buf(chunk);
    try
    {
      kmd_host::Lz4FrameReader r(in);
      for (size_t got; (got = r.read(buf.data(), chunk)) > 0;) fwrite(buf.data(), 1, got, out);
    }
    catch (const std::exception& e) { fprintf(stderr, "%s\n", e.what()); return 3; }
    fclose(in); fclose(out);
    return 0;
  }
  if (argc == 4 && !strcmp(argv[1], "encode"))
  {
    FILE* in = fopen(argv[2], "rb");
    FILE* out = fopen(argv[3], "wb");
    if (!in || !out) return 2;
    {
      kmd_host::Lz4FrameWriter w(out);
      std::vector<uint8_t> buf(100003);
      for (size_t got; (got = fread(buf.data(), 1, buf.size(), in)) > 0;) w.write(buf.data(), got);
      w.close();
    }
    fclose(in); fclose(out);
    return 0;
  }
  return 2;
}
