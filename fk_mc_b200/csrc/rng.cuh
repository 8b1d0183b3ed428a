// Device-side std::mt19937 and the libstdc++ (GCC >= 11) distributions the reference draws from,
// bit-exact so that a chain consumes the same random stream as the reference's
//   std::mt19937 random(SEED + rank)                        src/mc_metropolis.cpp:25
//   std::uniform_int_distribution<>(0, V-1)                 src/moves.cpp:7,55; src/configuration.cpp:48
//   std::uniform_real_distribution<>(0,1)                   include/fk_mc/mc_metropolis.hpp:143
//   std::discrete_distribution<>                            include/fk_mc/mc_metropolis.hpp:141
// State per chain: 624 words + read index, in global memory.
#pragma once
#include <stdint.h>

#define FKMC_MT_N 624
#define FKMC_MT_WORDS 625

struct mt19937_dev {
    uint32_t* s;  // s[0..623] state, s[624] index
    __device__ __forceinline__ explicit mt19937_dev(uint32_t* state) : s(state) {}
    __device__ void seed(uint32_t value) {
        s[0] = value;
        for (int i = 1; i < FKMC_MT_N; ++i) s[i] = 1812433253u * (s[i - 1] ^ (s[i - 1] >> 30)) + (uint32_t)i;
        s[FKMC_MT_N] = FKMC_MT_N;
    }
    __device__ void twist() {
        const uint32_t UP = 0x80000000u, LO = 0x7fffffffu, MAG = 0x9908b0dfu;
        for (int k = 0; k < FKMC_MT_N - 397; ++k) {
            const uint32_t y = (s[k] & UP) | (s[k + 1] & LO);
            s[k] = s[k + 397] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
        }
        for (int k = FKMC_MT_N - 397; k < FKMC_MT_N - 1; ++k) {
            const uint32_t y = (s[k] & UP) | (s[k + 1] & LO);
            s[k] = s[k + (397 - FKMC_MT_N)] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
        }
        const uint32_t y = (s[FKMC_MT_N - 1] & UP) | (s[0] & LO);
        s[FKMC_MT_N - 1] = s[396] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
        s[FKMC_MT_N] = 0;
    }
    __device__ uint32_t next() {
        if (s[FKMC_MT_N] >= FKMC_MT_N) twist();
        uint32_t y = s[s[FKMC_MT_N]++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    // std::uniform_int_distribution<int>(0, range-1): Lemire's nearly-divisionless method
    // (bits/uniform_int_dist.h _S_nd with a 64-bit product)
    __device__ uint32_t uniform_int(uint32_t range) {
        uint64_t product = (uint64_t)next() * (uint64_t)range;
        uint32_t low = (uint32_t)product;
        if (low < range) {
            const uint32_t threshold = (0u - range) % range;
            while (low < threshold) {
                product = (uint64_t)next() * (uint64_t)range;
                low = (uint32_t)product;
            }
        }
        return (uint32_t)(product >> 32);
    }
    // std::generate_canonical<double, 53>: two words, low word first (bits/random.tcc)
    __device__ double canonical() {
        const double w1 = (double)next();
        const double w2 = (double)next();
        double r = (w1 + w2 * 4294967296.0) / 18446744073709551616.0;
        if (r >= 1.0) r = 0.99999999999999988897769753748434595763683319091796875;
        return r;
    }
};
