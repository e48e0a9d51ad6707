// Host-side mirror of the reference's AndersonAcceleration classes over the C ABI
// (include/aaadmm.h). Same member names, argument order and meaning; all state and all
// arithmetic live on the GPU.
//
//   AndersonAccelerationH  <-  admm_anderson_hard_zxu/src/AndersonAcceleration.h:38-212
//                              (byte-identical to Geometry/AndersonAcceleration.h)
//   AndersonAccelerationX  <-  admm_anderson_xzu/src/AndersonAcceleration.h:39-295
//
// The reference templates over Eigen::PlainObjectBase<Derived>; these templates accept any
// contiguous container with data() and size() (Eigen vectors/matrices, std::vector<double>).
// Define AAADMM_ANDERSON_ALIAS_H or AAADMM_ANDERSON_ALIAS_X before including to get the
// reference's class name `AndersonAcceleration`.
#pragma once
#include <cassert>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/aaadmm.h"

namespace aaadmm {

class AndersonAccelerationH {
public:
    AndersonAccelerationH(int m, int total_dim, int effective_dim)
        : m_(m), total_dim_(total_dim), effective_dim_(effective_dim) {
        assert(m_ > 0);
        if (aaadmm_aa_create(&h_, m, total_dim, effective_dim) != 0)
            throw std::runtime_error(std::string("AndersonAcceleration: ") + aaadmm_last_error());
        buf_.resize(total_dim);
        out_.resize(total_dim);
    }
    ~AndersonAccelerationH() { aaadmm_aa_destroy(h_); }
    AndersonAccelerationH(const AndersonAccelerationH &) = delete;
    AndersonAccelerationH &operator=(const AndersonAccelerationH &) = delete;

    template <typename V>
    void replace(const V &u) {
        check(aaadmm_aa_replace(h_, u.data(), (int64_t)u.size()));
    }
    // The first argument must be the effective variable
    template <typename V1, typename V2>
    void replace(const V1 &u1, const V2 &u2) {
        pack(u1, u2);
        check(aaadmm_aa_replace(h_, buf_.data(), total_dim_));
    }
    template <typename V>
    void reset(const V &u) {
        check(aaadmm_aa_reset(h_, u.data(), (int64_t)u.size()));
    }
    template <typename V1, typename V2>
    void reset(const V1 &u1, const V2 &u2) {
        pack(u1, u2);
        check(aaadmm_aa_reset(h_, buf_.data(), total_dim_));
    }
    template <typename V>
    void compute(const V &g, V &accel_u) {
        check(aaadmm_aa_compute(h_, g.data(), accel_u.data(), (int64_t)g.size()));
    }
    template <typename V1, typename V2>
    void compute(const V1 &g1, const V2 &g2, V1 &accel_u1, V2 &accel_u2) {
        pack(g1, g2);
        check(aaadmm_aa_compute(h_, buf_.data(), out_.data(), total_dim_));
        std::copy(out_.begin(), out_.begin() + accel_u1.size(), accel_u1.data());
        std::copy(out_.begin() + accel_u1.size(), out_.begin() + accel_u1.size() + accel_u2.size(), accel_u2.data());
    }
    template <typename V>
    void init(const V &init_u) {
        assert(int(init_u.size()) == total_dim_);
        check(aaadmm_aa_init(h_, init_u.data(), (int64_t)init_u.size()));
    }
    template <typename V1, typename V2>
    void init(const V1 &u1, const V2 &u2) {
        assert(int(u1.size() + u2.size()) == total_dim_);
        pack(u1, u2);
        check(aaadmm_aa_init(h_, buf_.data(), total_dim_));
    }

private:
    int m_, total_dim_, effective_dim_;
    aaadmm_aa *h_ = nullptr;
    std::vector<double> buf_, out_;
    template <typename V1, typename V2>
    void pack(const V1 &a, const V2 &b) {
        std::copy(a.data(), a.data() + a.size(), buf_.begin());
        std::copy(b.data(), b.data() + b.size(), buf_.begin() + a.size());
    }
    static void check(int rc) {
        if (rc != 0) throw std::runtime_error(std::string("AndersonAcceleration: ") + aaadmm_last_error());
    }
};

class AndersonAccelerationX {
public:
    AndersonAccelerationX() {}
    ~AndersonAccelerationX() { aaadmm_aa_destroy(h_); }
    AndersonAccelerationX(const AndersonAccelerationX &) = delete;
    AndersonAccelerationX &operator=(const AndersonAccelerationX &) = delete;

    template <typename V>
    void replace(const V &g) {
        check(aaadmm_aa_replace(h_, g.data(), (int64_t)g.size()));
    }
    template <typename V>
    void compute(V &curr_g, const V &g) {
        assert(h_);
        check(aaadmm_aa_compute(h_, g.data(), curr_g.data(), (int64_t)g.size()));
    }
    // m: number of previous iterations used; d: dimension of variables; g0: initial values
    template <typename V>
    void init(int m, int d, const V &g0) {
        assert(m > 0);
        if (h_) aaadmm_aa_destroy(h_);
        h_ = nullptr;
        check(aaadmm_aa_create(&h_, m, d, d));
        check(aaadmm_aa_init(h_, g0.data(), d));
    }

private:
    aaadmm_aa *h_ = nullptr;
    static void check(int rc) {
        if (rc != 0) throw std::runtime_error(std::string("AndersonAcceleration: ") + aaadmm_last_error());
    }
};

}  // namespace aaadmm

#if defined(AAADMM_ANDERSON_ALIAS_H)
typedef aaadmm::AndersonAccelerationH AndersonAcceleration;
#elif defined(AAADMM_ANDERSON_ALIAS_X)
typedef aaadmm::AndersonAccelerationX AndersonAcceleration;
#endif
