// Host random numbers of the reference's examples: `rand_chacha::ChaCha20Rng::seed_from_u64(seed)`
// feeds Environment::reset_parameter (environment.rs:16-40,190-202: Open01 samples through Box-Muller / an affine map), the
// per-epoch shuffle (examples/fashion_mnist/main.rs:382), the random pixel batches (examples/image_fit/main.rs:377-378) and
// the per-step rand_seed (`rng.next_u32()`).  The crates are not in the reference tree (Cargo.toml: rand = "0.8",
// rand_chacha = "0.3"; Cargo.lock pins rand 0.8.x / rand_core 0.6.x / rand_chacha 0.3.x), so this restates their published
// algorithms:
//   * rand_core 0.6 `SeedableRng::seed_from_u64`: a PCG32 stream (multiplier 6364136223846793005, increment
//     11634580027462260723, xorshift 18 / 27, rotate by the top 5 bits) fills the 32-byte seed four bytes at a time;
//   * rand_chacha 0.3 `ChaCha20Rng`: the ChaCha block function with 20 rounds, key = seed, 64-bit block counter in words
//     12-13 starting at 0, stream id 0 in words 14-15; results are consumed in order through a 64-word buffer
//     (rand_core `BlockRng`: next_u64 takes two consecutive words, low word first, straddling a refill when one is left);
//   * rand 0.8 `Open01` for f32: 23 high bits of one u32 as the fraction of a float in [1, 2), minus (1 - EPSILON / 2);
//   * rand 0.8 `gen_range` on integers (`UniformInt::sample_single_inclusive`): widening multiply with the rejection zone
//     (range << leading_zeros) - 1; u32 ranges draw u32 words, usize ranges draw u64;
//   * rand 0.8 `SliceRandom::shuffle`: for i = len - 1 down to 1 swap(i, gen_index(i + 1)), gen_index through the u32
//     range when the bound fits.
// Pinned by the ChaCha20 block known answer (all-zero key: 0xade0b876, 0x903df1a0, ... -- rand_chacha's own
// test_chacha_true_values_a and the zero-key / zero-nonce vector of the ChaCha reference) in tests/test_cpu_host_io.py; the
// seed expansion and the sampling rules are restated from memory of the crates' sources and have no vector here: parity with
// the crates beyond the block function is UNPINNED.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <utility>

namespace descent {

class ChaCha20Rng {
public:
    static ChaCha20Rng seed_from_u64(uint64_t state) {
        uint8_t seed[32];
        for (int i = 0; i < 8; ++i) {
            state = state * 6364136223846793005ull + 11634580027462260723ull;
            const uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
            const uint32_t rot = (uint32_t)(state >> 59);
            const uint32_t x = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
            seed[4 * i] = (uint8_t)x; seed[4 * i + 1] = (uint8_t)(x >> 8); seed[4 * i + 2] = (uint8_t)(x >> 16); seed[4 * i + 3] = (uint8_t)(x >> 24);
        }
        return from_seed(seed);
    }
    static ChaCha20Rng from_seed(const uint8_t seed[32]) {
        ChaCha20Rng r;
        for (int i = 0; i < 8; ++i)
            r.key_[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) | ((uint32_t)seed[4 * i + 3] << 24);
        r.counter_ = 0;
        r.index_ = 64;  // empty buffer
        return r;
    }
    uint32_t next_u32() {
        if (index_ >= 64) refill(0);
        return buffer_[index_++];
    }
    uint64_t next_u64() {
        if (index_ < 63) {
            const uint64_t v = ((uint64_t)buffer_[index_ + 1] << 32) | buffer_[index_];
            index_ += 2;
            return v;
        }
        if (index_ >= 64) {
            refill(2);
            return ((uint64_t)buffer_[1] << 32) | buffer_[0];
        }
        const uint64_t lo = buffer_[63];
        refill(1);
        return ((uint64_t)buffer_[0] << 32) | lo;
    }
    float open01() {
        const uint32_t fraction = next_u32() >> 9;
        const uint32_t bits = fraction | 0x3f800000u;
        float f;
        std::memcpy(&f, &bits, 4);
        return f - (1.0f - 1.1920928955078125e-7f / 2.0f);
    }
    // uniform in [low, high), high > low
    uint32_t gen_range_u32(uint32_t low, uint32_t high) {
        const uint32_t range = high - 1 - low + 1;
        if (range == 0) return next_u32();
        const uint32_t zone = (range << __builtin_clz(range)) - 1;
        for (;;) {
            const uint64_t m = (uint64_t)next_u32() * range;
            if ((uint32_t)m <= zone) return low + (uint32_t)(m >> 32);
        }
    }
    uint64_t gen_range_u64(uint64_t low, uint64_t high) {
        const uint64_t range = high - 1 - low + 1;
        if (range == 0) return next_u64();
        const uint64_t zone = (range << __builtin_clzll(range)) - 1;
        for (;;) {
            const unsigned __int128 m = (unsigned __int128)next_u64() * range;
            if ((uint64_t)m <= zone) return low + (uint64_t)(m >> 64);
        }
    }
    template <class T>
    void shuffle(T* data, size_t len) {
        for (size_t i = len; i-- > 1;) {
            const size_t bound = i + 1;
            const size_t j = bound <= 0xffffffffull ? (size_t)gen_range_u32(0, (uint32_t)bound) : (size_t)gen_range_u64(0, bound);
            std::swap(data[i], data[j]);
        }
    }

private:
    static uint32_t rotl(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
    static void quarter(uint32_t* x, int a, int b, int c, int d) {
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16);
        x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);
        x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
    }
    void block(uint32_t* out) {
        uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key_[0], key_[1], key_[2], key_[3], key_[4], key_[5], key_[6], key_[7],
                           (uint32_t)counter_, (uint32_t)(counter_ >> 32), 0u, 0u};
        uint32_t x[16];
        std::memcpy(x, in, sizeof(x));
        for (int round = 0; round < 10; ++round) {
            quarter(x, 0, 4, 8, 12); quarter(x, 1, 5, 9, 13); quarter(x, 2, 6, 10, 14); quarter(x, 3, 7, 11, 15);
            quarter(x, 0, 5, 10, 15); quarter(x, 1, 6, 11, 12); quarter(x, 2, 7, 8, 13); quarter(x, 3, 4, 9, 14);
        }
        for (int i = 0; i < 16; ++i) out[i] = x[i] + in[i];
        counter_ += 1;
    }
    void refill(size_t index) {
        for (int b = 0; b < 4; ++b) block(buffer_ + 16 * b);
        index_ = index;
    }
    uint32_t key_[8];
    uint64_t counter_ = 0;
    uint32_t buffer_[64];
    size_t index_ = 64;
};

}  // namespace descent
