// monopsr_b200/csrc/c_abi.cu -- library identification and launch accounting.
#include "common.cuh"
#include "../../include/monopsr_b200_tfops.h"

namespace mpb {
unsigned long long g_launch_count = 0;
}

MPB_API const char* mpb_version(void) { return "monopsr_b200 0.1 sm_100a"; }

MPB_API unsigned long long mpb_launch_count(void) {
    return __atomic_load_n(&mpb::g_launch_count, __ATOMIC_RELAXED);
}

// CRC-32C (Castagnoli, reflected polynomial 0x82F63B78), slicing-by-8 on the host: the checksum TensorFlow stores for
// every tensor of a checkpoint (core/tf_checkpoint.py).  crc = previous value (0 to start); plain host code.
namespace mpb {
static unsigned g_crc_tab[8][256];
static bool g_crc_ready = false;
static void crc_init() {
    for (unsigned i = 0; i < 256; i++) {
        unsigned c = i;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
        g_crc_tab[0][i] = c;
    }
    for (unsigned i = 0; i < 256; i++)
        for (int t = 1; t < 8; t++) g_crc_tab[t][i] = (g_crc_tab[t - 1][i] >> 8) ^ g_crc_tab[0][g_crc_tab[t - 1][i] & 0xFFu];
    g_crc_ready = true;
}
}  // namespace mpb

MPB_API unsigned mpb_crc32c(const void* data, unsigned long long n, unsigned crc) {
    using namespace mpb;
    if (!g_crc_ready) crc_init();
    const unsigned char* p = static_cast<const unsigned char*>(data);
    unsigned c = crc ^ 0xFFFFFFFFu;
    while (n >= 8) {
        unsigned lo, hi;
        __builtin_memcpy(&lo, p, 4);
        __builtin_memcpy(&hi, p + 4, 4);
        lo ^= c;
        c = g_crc_tab[7][lo & 0xFFu] ^ g_crc_tab[6][(lo >> 8) & 0xFFu] ^ g_crc_tab[5][(lo >> 16) & 0xFFu] ^
            g_crc_tab[4][lo >> 24] ^ g_crc_tab[3][hi & 0xFFu] ^ g_crc_tab[2][(hi >> 8) & 0xFFu] ^
            g_crc_tab[1][(hi >> 16) & 0xFFu] ^ g_crc_tab[0][hi >> 24];
        p += 8;
        n -= 8;
    }
    while (n--) c = g_crc_tab[0][(c ^ *p++) & 0xFFu] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}
