// Host build of the multi-buffer MD5 used by the host->host encode path (pyflac_b200/csrc/md5_host.h, md5_mb.h) so it can
// be checked against hashlib on CPU-only machines (tests/test_abi_cpu.py).
#include <cstddef>
#include <cstdint>
#include <cstring>
#include "../../pyflac_b200/csrc/md5_host.h"
#include "../../pyflac_b200/csrc/md5_mb.h"
extern "C" {
int t_md5_lanes() { return fb::md5_mb16_available() ? 16 : (fb::md5_mb_available() ? 8 : 1); }
// digests of n (<= 16) byte strings through the widest kernel the CPU has; lanes: 16, 8 or 1 (scalar) forces a path
int t_md5_group(const uint8_t* blob, const uint64_t* off, const uint64_t* len, int n, int lanes, uint8_t* digests) {
    const uint8_t* d[16]; size_t l[16];
    for (int i = 0; i < 16; i++) { d[i] = blob + off[i < n ? i : 0]; l[i] = (size_t)len[i < n ? i : 0]; }
    if (lanes == 16) {
        if (!fb::md5_mb16_available()) return -1;
        uint8_t dig[16][16]; fb::md5_group16(d, l, n, dig); memcpy(digests, dig, (size_t)n * 16); return 0;
    }
    if (lanes == 8) {
        if (!fb::md5_mb_available() || n > 8) return -1;
        uint8_t dig[8][16]; fb::md5_group8(d, l, n, dig); memcpy(digests, dig, (size_t)n * 16); return 0;
    }
    for (int i = 0; i < n; i++) { fb::Md5 m; m.init(); m.update(d[i], l[i]); m.final(digests + 16 * i); }
    return 0;
}
}
