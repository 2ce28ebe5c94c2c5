"""Pure-Python restatement of the host random numbers the reference's examples draw (TEST INFRASTRUCTURE ONLY):
rand_chacha 0.3 `ChaCha20Rng`, rand_core 0.6 `seed_from_u64`, rand 0.8 `Open01` / `gen_range` / `shuffle`
(examples/fashion_mnist/main.rs:362-386, examples/image_fit/main.rs:352-395, src/environment.rs:16-40).  The crates are
not vendored in the reference tree; the ChaCha block function is pinned by its published known answer
(tests/test_cpu_host_io.py), the seed expansion and sampling rules are restated from the crates' published sources and
are otherwise "parity unpinned"."""
import struct

M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF


def _rotl(v, n):
    return ((v << n) | (v >> (32 - n))) & M32


def chacha20_block(key_words, counter, stream=0):
    """16 output words of the ChaCha20 block function: constants, 8 key words, 64-bit counter, 64-bit stream id."""
    state = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + [counter & M32, (counter >> 32) & M32, stream & M32, (stream >> 32) & M32]
    x = list(state)

    def quarter(a, b, c, d):
        x[a] = (x[a] + x[b]) & M32; x[d] = _rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & M32; x[b] = _rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & M32; x[d] = _rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & M32; x[b] = _rotl(x[b] ^ x[c], 7)

    for _ in range(10):
        quarter(0, 4, 8, 12); quarter(1, 5, 9, 13); quarter(2, 6, 10, 14); quarter(3, 7, 11, 15)
        quarter(0, 5, 10, 15); quarter(1, 6, 11, 12); quarter(2, 7, 8, 13); quarter(3, 4, 9, 14)
    return [(a + b) & M32 for a, b in zip(x, state)]


class ChaCha20Rng:
    def __init__(self, seed_bytes):
        self.key = list(struct.unpack("<8I", bytes(seed_bytes)))
        self.counter = 0
        self.buffer = []
        self.index = 64

    @classmethod
    def seed_from_u64(cls, state):
        seed = b""
        for _ in range(8):  # rand_core 0.6: a PCG32 stream fills the seed four bytes at a time
            state = (state * 6364136223846793005 + 11634580027462260723) & M64
            xorshifted = (((state >> 18) ^ state) >> 27) & M32
            rot = state >> 59
            x = ((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & M32
            seed += struct.pack("<I", x)
        return cls(seed)

    def _refill(self, index):
        self.buffer = []
        for _ in range(4):
            self.buffer += chacha20_block(self.key, self.counter)
            self.counter += 1
        self.index = index

    def next_u32(self):
        if self.index >= 64:
            self._refill(0)
        v = self.buffer[self.index]
        self.index += 1
        return v

    def next_u64(self):  # rand_core BlockRng::next_u64
        if self.index < 63:
            v = (self.buffer[self.index + 1] << 32) | self.buffer[self.index]
            self.index += 2
            return v
        if self.index >= 64:
            self._refill(2)
            return (self.buffer[1] << 32) | self.buffer[0]
        lo = self.buffer[63]
        self._refill(1)
        return (self.buffer[0] << 32) | lo

    def open01(self):
        import numpy as np
        bits = (self.next_u32() >> 9) | 0x3F800000
        return np.float32(struct.unpack("<f", struct.pack("<I", bits))[0]) - (np.float32(1.0) - np.float32(2.0 ** -23) / np.float32(2.0))

    def gen_range(self, low, high, u32=False):
        bits = 32 if u32 else 64
        mask = (1 << bits) - 1
        rng_range = (high - 1 - low + 1) & mask
        if rng_range == 0:
            return self.next_u32() if u32 else self.next_u64()
        zone = ((rng_range << (bits - rng_range.bit_length())) - 1) & mask
        while True:
            v = self.next_u32() if u32 else self.next_u64()
            m = v * rng_range
            if (m & mask) <= zone:
                return low + (m >> bits)

    def shuffle(self, items):
        items = list(items)
        for i in range(len(items) - 1, 0, -1):
            j = self.gen_range(0, i + 1, u32=(i + 1) <= M32)
            items[i], items[j] = items[j], items[i]
        return items
