// Version / error reporting of the C ABI (include/pwc_b200.h).
#include "common.cuh"
#include <cstdarg>
#include <cstring>

namespace pwc {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
// Number of SMs of the current device (persistent grids are sized from it; cached per device).
int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}
}  // namespace pwc

extern "C" int pwc_version(void) { return PWC_ABI_VERSION; }
extern "C" const char* pwc_last_error(void) { return pwc::g_err; }

// CRC-32C, slice-by-8 (host).  Tables are built on first use.
extern "C" unsigned int pwc_crc32c(const void* data, long long n, unsigned int crc) {
    static uint32_t tab[8][256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
            tab[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; ++i)
            for (int t = 1; t < 8; ++t) tab[t][i] = (tab[t - 1][i] >> 8) ^ tab[0][tab[t - 1][i] & 0xFF];
        ready = true;
    }
    const uint8_t* p = static_cast<const uint8_t*>(data);
    uint32_t c = crc ^ 0xFFFFFFFFu;
    while (n > 0 && (reinterpret_cast<uintptr_t>(p) & 7)) { c = tab[0][(c ^ *p++) & 0xFF] ^ (c >> 8); --n; }
    while (n >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        w ^= c;
        c = tab[7][w & 0xFF] ^ tab[6][(w >> 8) & 0xFF] ^ tab[5][(w >> 16) & 0xFF] ^ tab[4][(w >> 24) & 0xFF] ^
            tab[3][(w >> 32) & 0xFF] ^ tab[2][(w >> 40) & 0xFF] ^ tab[1][(w >> 48) & 0xFF] ^ tab[0][(w >> 56) & 0xFF];
        p += 8; n -= 8;
    }
    while (n-- > 0) c = tab[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}
