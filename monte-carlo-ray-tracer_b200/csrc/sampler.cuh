// Hash-based Owen-scrambled Sobol sampler, bit-exact with the reference's
// Sampler (source/sampling/sampler.hpp:13-91) and Sobol (source/sampling/sobol.hpp:7-72).
//
// The reference keeps five uint32 in thread_local storage; on the device the whole state is a pure
// function of (global_seed, pixel, sample, number of shuffle() calls), so a path carries only
// (pixel, sample, depth) and rebuilds the registers each bounce:
//     base_seed      = hashCombine(global_seed, hash(pixel))            initiate()   sampler.hpp:30-33
//     bit_reversed   = reverseBits(sample)                               setIndex()   :36-42
//     seed           = hashCombine(base_seed, hash(sequence))            shuffle()    :46-50
//     shuffled_index = scramble(bit_reversed, seed)
//     get<D>()       = scramble(sobol<D>(shuffled_index), hashCombine(seed, hash(D))) * 2^-32   :19-27
// The 6x32 direction numbers are generated at compile time from the Joe-Kuo primitive polynomials
// (s, a, m_i) of dimensions 2..7 — the published data the reference cites (sobol.hpp:19-31) — and
// stored bit-reversed; with fully unrolled loops they become LOP3 immediates, and the per-bit lane
// masks are shared by all dimensions requested in one call.
#pragma once

#include "vec.cuh"

namespace mcrt
{
    MCRT_HD constexpr uint32_t reverseBits32(uint32_t x)
    {
        x = ((x & 0xaaaaaaaau) >> 1) | ((x & 0x55555555u) << 1);
        x = ((x & 0xccccccccu) >> 2) | ((x & 0x33333333u) << 2);
        x = ((x & 0xf0f0f0f0u) >> 4) | ((x & 0x0f0f0f0fu) << 4);
        x = ((x & 0xff00ff00u) >> 8) | ((x & 0x00ff00ffu) << 8);
        return (x >> 16) | (x << 16);
    }

    struct SobolTable { uint32_t v[6][32]; };

    // Joe & Kuo "new-joe-kuo-6.21201", dimensions 2-7: degree s, coefficient bits a, initial m_i.
    MCRT_HD constexpr SobolTable makeSobolTable()
    {
        const uint32_t deg[6] = { 1, 2, 3, 3, 4, 4 };
        const uint32_t poly[6] = { 0, 1, 1, 2, 1, 4 };
        const uint32_t init[6][4] = { { 1, 0, 0, 0 }, { 1, 3, 0, 0 }, { 1, 3, 1, 0 },
                                      { 1, 1, 1, 0 }, { 1, 1, 3, 3 }, { 1, 3, 5, 13 } };
        SobolTable t{};
        for (uint32_t d = 0; d < 6; d++)
        {
            const uint32_t s = deg[d];
            uint32_t dir[32] = {};
            for (uint32_t b = 0; b < 32; b++)
            {
                if (b < s)
                {
                    dir[b] = init[d][b] << (31 - b);
                }
                else
                {
                    uint32_t v = dir[b - s] ^ (dir[b - s] >> s);
                    for (uint32_t k = 1; k < s; k++)
                    {
                        if ((poly[d] >> (s - 1 - k)) & 1u) v ^= dir[b - k];
                    }
                    dir[b] = v;
                }
            }
            for (uint32_t b = 0; b < 32; b++) t.v[d][b] = reverseBits32(dir[b]);
        }
        return t;
    }

    // hash-prospector 2-round hash, sampler.hpp:78-86
    MCRT_HD constexpr uint32_t samplerHash(uint32_t x)
    {
        x ^= x >> 15; x *= 0xd168aaadu;
        x ^= x >> 15; x *= 0xaf723597u;
        x ^= x >> 15;
        return x;
    }

    // boost-style combine, sampler.hpp:89-92
    MCRT_HD constexpr uint32_t samplerHashCombine(uint32_t seed, uint32_t v)
    {
        return seed ^ (v + 0x9e3779b9u + (seed << 6) + (seed >> 2));
    }

    // Laine-Karras style scramble in the bit-reversed domain, sampler.hpp:61-73
    MCRT_HD constexpr uint32_t samplerScramble(uint32_t x, uint32_t seed)
    {
        x ^= x * 0x3d20adeau;
        x += seed;
        x *= (seed >> 16) | 1u;
        x ^= x * 0x05526c56u;
        x ^= x * 0x53a22864u;
        return reverseBits32(x);
    }

    // Dimension layout, source/sampling/sampling.hpp:59-76
    enum SampleDim
    {
        DIM_PIXEL = 0, DIM_LENS = 2,
        DIM_LIGHT = 0, DIM_BSDF = 3, DIM_INTERACTION = 5, DIM_ABSORB = 6,
        DIM_PM_LIGHT = 0, DIM_PM_REJECT = 2
    };

    struct SobolByteTables;

    struct SamplerState
    {
        uint32_t seed;
        uint32_t shuffled_index;
        const SobolByteTables* tab;   // shared-memory byte tables (null: bit loop)

        // state after initiate(pixel), setIndex(sample) and `sequence` shuffle() calls
        MCRT_HD static SamplerState make(uint32_t global_seed, uint32_t pixel, uint32_t sample, uint32_t sequence)
        {
            SamplerState s;
            s.tab = nullptr;
            uint32_t base_seed = samplerHashCombine(global_seed, samplerHash(pixel));
            if (sequence == 0)
            {
                s.seed = base_seed;
                s.shuffled_index = sample;
            }
            else
            {
                s.seed = samplerHashCombine(base_seed, samplerHash(sequence));
                s.shuffled_index = samplerScramble(reverseBits32(sample), s.seed);
            }
            return s;
        }

        // Raw 32-bit values of the dimensions selected by MASK (bit d = dimension d), written to
        // out[d]. All loops unroll; unused dimensions cost nothing.
        template <uint32_t MASK>
        MCRT_HD void raw(uint32_t* out) const
        {
            constexpr SobolTable T = makeSobolTable();
            uint32_t acc[7] = { shuffled_index, 0u, 0u, 0u, 0u, 0u, 0u };
            if constexpr ((MASK & ~1u) != 0u)
            {
#pragma unroll
                for (int b = 0; b < 32; b++)
                {
                    const uint32_t lane = 0u - ((shuffled_index >> b) & 1u);
#pragma unroll
                    for (int d = 1; d < 7; d++)
                    {
                        if ((MASK >> d) & 1u) acc[d] ^= lane & T.v[d - 1][b];
                    }
                }
            }
#pragma unroll
            for (int d = 0; d < 7; d++)
            {
                if ((MASK >> d) & 1u) out[d] = samplerScramble(acc[d], samplerHashCombine(seed, samplerHash((uint32_t)d)));
            }
        }
    };

    // x * 2^-32 (sampler.hpp:26). Exact in double; in float the product can round up to 1.0, which
    // the reference's double can never produce, so the float path clamps to the largest value < 1.
    MCRT_HD double unitFromBits(uint32_t x, double) { return (double)x * 0x1p-32; }
    MCRT_HD float unitFromBits(uint32_t x, float)
    {
        float f = (float)x * 0x1p-32f;
        return f < 1.0f ? f : 0x1.fffffep-1f;
    }

    // Byte-sliced Sobol matrices for kernels that draw many dimensions per thread: entry [d][j][v] is
    // the XOR of the direction numbers of dimension d+1 selected by byte j = v of the index, so one
    // dimension costs 4 shared-memory lookups + 3 XOR instead of a 32-step bit loop. 6 x 4 x 256 x 4 B
    // = 24 KB, filled per CTA from the compile-time direction numbers.
    struct SobolByteTables
    {
        uint32_t t[6][4][256];

        // cooperative copy of the host-built tables (makeSobolByteTable) into shared memory
        MCRT_D void fill(const uint32_t* global_tables)
        {
            uint32_t* flat = &t[0][0][0];
            for (uint32_t e = threadIdx.x; e < 6u * 4u * 256u; e += blockDim.x) flat[e] = __ldg(&global_tables[e]);
        }

        // raw 32-bit value of dimension `dim` (0..6) for the sampler state
        MCRT_D uint32_t raw(const SamplerState& s, uint32_t dim) const
        {
            uint32_t acc = s.shuffled_index;
            if (dim > 0)
            {
                const uint32_t i = s.shuffled_index;
                acc = t[dim - 1][0][i & 255u] ^ t[dim - 1][1][(i >> 8) & 255u] ^ t[dim - 1][2][(i >> 16) & 255u] ^ t[dim - 1][3][i >> 24];
            }
            return samplerScramble(acc, samplerHashCombine(s.seed, samplerHash(dim)));
        }

        template <class R, int START, int N>
        MCRT_D void get(const SamplerState& s, R* u) const
        {
#pragma unroll
            for (int k = 0; k < N; k++) u[k] = unitFromBits(raw(s, (uint32_t)(START + k)), R(0));
        }
    };

    // host-side builder of the byte tables (uploaded once per context)
    inline void makeSobolByteTable(uint32_t* out /* [6*4*256] */)
    {
        const SobolTable T = makeSobolTable();
        for (uint32_t d = 0; d < 6; d++)
            for (uint32_t j = 0; j < 4; j++)
                for (uint32_t v = 0; v < 256; v++)
                {
                    uint32_t x = 0;
                    for (uint32_t b = 0; b < 8; b++) if ((v >> b) & 1u) x ^= T.v[d][j * 8 + b];
                    out[(d * 4 + j) * 256 + v] = x;
                }
    }

    template <class R, int START, int N>
    MCRT_HD void samplerGet(const SamplerState& s, R* u)
    {
#ifdef __CUDA_ARCH__
        if (s.tab) { s.tab->template get<R, START, N>(s, u); return; }
#endif
        uint32_t raw[7];
        s.raw<(((1u << N) - 1u) << START)>(raw);
#pragma unroll
        for (int i = 0; i < N; i++) u[i] = unitFromBits(raw[START + i], R(0));
    }
}
