// ORACLE / TEST INFRASTRUCTURE — never linked into the product library.
//
// Force-included (-include) in front of every reference translation unit when oracle/_ref is built.
// The reference seeds its sampler with std::random_device{}() (source/sampling/sampler.hpp:58) and
// is therefore non-deterministic. Rather than editing or copying the reference sources, this header
// pulls <random> in first and then renames the identifier, so the reference's own line becomes
// `std::mcrt_pinned_device{}()` and returns a fixed seed. The same rename hits the mt19937_64 engine
// that only shuffles bucket/work order (source/sampling/sampling.hpp:50).
#pragma once

#include <random>
#include <cstdint>

namespace mcrt_oracle
{
    // MCRT_ORACLE_SEED from the environment, default 0x12345678. Defined in ref_driver.cpp.
    unsigned pinnedSeed();
}

namespace std
{
    struct mcrt_pinned_device
    {
        using result_type = unsigned;
        mcrt_pinned_device() { }
        result_type operator()() { return mcrt_oracle::pinnedSeed(); }
        static constexpr result_type min() { return 0u; }
        static constexpr result_type max() { return 0xFFFFFFFFu; }
    };
}

#define random_device mcrt_pinned_device
