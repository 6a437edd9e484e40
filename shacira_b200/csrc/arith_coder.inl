// arith_coder.inl -- host-side static arithmetic coder for the rounded latents (the "real
// bitstream" behind LatentGrid.size(use_torchac=True)).
//
// The reference only measures len(torchac.encode_float_cdf(...)) * 8 (latent_grid.py:155-172);
// torchac is an unvendored, unpinned, absent dependency, so byte parity with it is UNPINNED
// (DESIGN.md). This is a self-contained coder of the same family -- 32-bit low/high interval,
// 16-bit cumulative frequencies, pending-bit (underflow) handling, MSB-first bit packing, after
// Witten/Neal/Cleary and M. Nelson's "Data Compression With Arithmetic Coding" -- with its own
// decoder so that the stream is verifiable by round trip.
#include <cstdint>
#include <vector>

namespace shacira_ac {

constexpr int kPrecision = 16;
constexpr uint32_t kHalf = 0x80000000u, kQuarter = 0x40000000u, kThreeQuarter = 0xC0000000u;

struct BitWriter {
    uint8_t* out;
    int64_t cap, pos = 0;
    uint32_t acc = 0;
    int nbits = 0;
    bool overflow = false;
    void put(int bit) {
        acc = (acc << 1) | (uint32_t)bit;
        if (++nbits == 8) {
            if (pos < cap) out[pos] = (uint8_t)acc; else overflow = true;
            ++pos;
            acc = 0;
            nbits = 0;
        }
    }
    void put_with_pending(int bit, int64_t& pending) {
        put(bit);
        for (; pending > 0; --pending) put(!bit);
    }
    void flush() {
        while (nbits != 0) put(0);
    }
};

struct BitReader {
    const uint8_t* in;
    int64_t nbytes, pos = 0;
    uint32_t acc = 0;
    int nbits = 0;
    int get() {
        if (nbits == 0) {
            acc = pos < nbytes ? in[pos] : 0;  // zero padding past the end
            ++pos;
            nbits = 8;
        }
        --nbits;
        return (acc >> nbits) & 1;
    }
};

// cdf: nsym+1 entries, cdf[0] = 0 < cdf[1] < ... < cdf[nsym] = 2^16.
inline int64_t encode(const int16_t* sym, int64_t n, const uint32_t* cdf, int32_t nsym, uint8_t* out, int64_t cap) {
    BitWriter bw{out, cap};
    uint32_t low = 0, high = 0xFFFFFFFFu;
    int64_t pending = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t s = sym[i];
        if (s < 0 || s >= nsym) return -1;
        const uint64_t span = (uint64_t)high - low + 1;
        high = low + (uint32_t)((span * cdf[s + 1]) >> kPrecision) - 1;
        low = low + (uint32_t)((span * cdf[s]) >> kPrecision);
        for (;;) {
            if (high < kHalf) {
                bw.put_with_pending(0, pending);
            } else if (low >= kHalf) {
                bw.put_with_pending(1, pending);
            } else if (low >= kQuarter && high < kThreeQuarter) {
                ++pending;
                low -= kQuarter;
                high -= kQuarter;
            } else {
                break;
            }
            low <<= 1;
            high = (high << 1) | 1u;
        }
    }
    ++pending;
    bw.put_with_pending(low < kQuarter ? 0 : 1, pending);
    bw.flush();
    return bw.overflow ? -2 - bw.pos : bw.pos;  // -2-needed on overflow
}

inline int decode(const uint8_t* in, int64_t nbytes, const uint32_t* cdf, int32_t nsym, int16_t* sym, int64_t n) {
    BitReader br{in, nbytes};
    uint32_t low = 0, high = 0xFFFFFFFFu, value = 0;
    for (int b = 0; b < 32; ++b) value = (value << 1) | (uint32_t)br.get();
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t span = (uint64_t)high - low + 1;
        const uint32_t target = (uint32_t)(((((uint64_t)(value - low) + 1) << kPrecision) - 1) / span);
        // largest s with cdf[s] <= target
        int32_t a = 0, b = nsym;
        while (b - a > 1) {
            const int32_t m = (a + b) >> 1;
            if (cdf[m] <= target) a = m; else b = m;
        }
        sym[i] = (int16_t)a;
        high = low + (uint32_t)((span * cdf[a + 1]) >> kPrecision) - 1;
        low = low + (uint32_t)((span * cdf[a]) >> kPrecision);
        for (;;) {
            if (high < kHalf) {
            } else if (low >= kHalf) {
                value -= kHalf; low -= kHalf; high -= kHalf;
            } else if (low >= kQuarter && high < kThreeQuarter) {
                value -= kQuarter; low -= kQuarter; high -= kQuarter;
            } else {
                break;
            }
            low <<= 1;
            high = (high << 1) | 1u;
            value = (value << 1) | (uint32_t)br.get();
        }
    }
    return 0;
}

}  // namespace shacira_ac
