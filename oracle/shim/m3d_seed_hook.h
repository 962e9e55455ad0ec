/*
 * m3d_seed_hook.h -- TEST INFRASTRUCTURE ONLY.  Force-included (-include) in front of every
 * reference translation unit built by `make -C oracle _ref`: the reference seeds its sampler from
 * std::random_device (include/misc3d/utils.h:74-77) and offers no way to pass a seed, so the token
 * `random_device` is re-pointed at a device that hands out m3dref_seed, m3dref_seed+1, ... (one value
 * per sampler construction = per FitModel = per segmentation round).  The sources stay unmodified.
 */
#pragma once
#include <random>
extern "C" unsigned int m3dref_seed;
namespace std {
struct m3d_seeded_device {
    unsigned int operator()() { return m3dref_seed++; }
};
}  // namespace std
#define random_device m3d_seeded_device
