// ORACLE (test infrastructure, never shipped).  CPU restatements of halo2_proofs `arithmetic.rs`
// (`best_multiexp`, `best_fft`, `eval_polynomial`, `kate_division`, `parallelize`) and of
// `poly/domain.rs` (EvaluationDomain).  halo2_proofs is an un-vendored git dependency
// (heliaxdev/halo2 branch `taiga`, /root/reference/taiga_halo2/Cargo.toml:14-15); algorithms are
// restated from the published zcash/halo2 0.3 lineage.
#pragma once
#include <algorithm>
#include <cmath>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include "field.hpp"

namespace orc {

inline int& num_threads() { static int n = std::max(1u, std::thread::hardware_concurrency()); return n; }

// A persistent pool of worker threads, the counterpart of the rayon pool behind halo2's `multicore` feature: the prover calls
// `parallelize` thousands of times per proof (every FFT stage), so spawning threads per call -- what this file used to do --
// cost more than the arithmetic on a 128-thread host and made the CPU baseline look slower than the algorithm is.
class Pool {
 public:
  static Pool& get() { static Pool p; return p; }
  // runs task(0) .. task(count - 1) on up to min(count, num_threads()) threads (the caller takes part); returns when all are done.
  // Only as many workers as the job can use are woken: on a 128-thread host a 4-thread prover must not wake 127 sleepers per FFT stage.
  void run(size_t count, const std::function<void(size_t)>& task) {
    if (count == 0) return;
    const size_t want = std::min(count - 1, (size_t)std::max(0, num_threads() - 1));
    if (want == 0 || in_task()) { for (size_t i = 0; i < count; ++i) task(i); return; }
    std::unique_lock<std::mutex> job_lock(job_mu_);   // one job at a time
    size_t have;
    { std::lock_guard<std::mutex> lk(mu_);
      while (workers_.size() < want) workers_.emplace_back([this] { loop(); });   // workers are created on demand, never more than a job asked for
      have = workers_.size();
      task_ = &task; count_ = count; next_.store(0); pending_.store(count); tickets_ = want; }
    if (want >= have) cv_.notify_all(); else for (size_t i = 0; i < want; ++i) cv_.notify_one();
    work();
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return pending_.load() == 0 && active_ == 0; });   // nobody is still polling this job's counter
    tickets_ = 0; task_ = nullptr;
  }
  ~Pool() {
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }

 private:
  Pool() {}
  static bool& in_task() { static thread_local bool f = false; return f; }
  void work() {
    in_task() = true;
    for (;;) {
      size_t i = next_.fetch_add(1);
      if (i >= count_) break;
      (*task_)(i);
      if (pending_.fetch_sub(1) == 1) { std::lock_guard<std::mutex> lk(mu_); done_cv_.notify_all(); }
    }
    in_task() = false;
  }
  void loop() {
    for (;;) {
      { std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || tickets_ > 0; });
        if (stop_) return;
        --tickets_; ++active_; }
      work();
      { std::lock_guard<std::mutex> lk(mu_); --active_; }
      done_cv_.notify_all();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_, job_mu_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(size_t)>* task_ = nullptr;
  size_t count_ = 0, tickets_ = 0;   // tickets_: workers still invited to the current job (guarded by mu_)
  std::atomic<size_t> next_{0}, pending_{0};
  int active_ = 0;   // workers inside work() (guarded by mu_)
  bool stop_ = false;
};

// halo2 `parallelize`: split [0,n) into contiguous chunks, one per thread (at least `grain` elements each)
inline void parallel_for(size_t n, const std::function<void(size_t, size_t)>& f, size_t grain = 128) {
  size_t nt = (size_t)num_threads();
  if (nt > n / grain) nt = n / grain;
  if (nt <= 1) { f(0, n); return; }
  const size_t chunk = (n + nt - 1) / nt, parts = (n + chunk - 1) / chunk;
  Pool::get().run(parts, [&](size_t t) { f(t * chunk, std::min(n, (t + 1) * chunk)); });
}

// ---- best_multiexp (bucket method, window c = ceil(ln n), no precomputation, per-thread chunks)
template <class F>
Jac<F> multiexp_serial(const u64 (*scalars)[4], const Affine<F>* bases, size_t n) {
  int c = n < 4 ? 1 : (n < 32 ? 3 : (int)std::ceil(std::log((double)n)));
  int segments = 256 / c + 1;
  Jac<F> acc = Jac<F>::identity();
  std::vector<Jac<F>> buckets((size_t(1) << c) - 1);
  for (int seg = segments - 1; seg >= 0; --seg) {
    for (int i = 0; i < c; ++i) acc = acc.dbl();
    for (auto& b : buckets) b = Jac<F>::identity();
    for (size_t i = 0; i < n; ++i) {
      // c bits of the canonical scalar starting at bit seg*c
      int skip = seg * c; if (skip >= 256) continue;
      int limb = skip / 64, off = skip % 64;
      u64 v = scalars[i][limb] >> off;
      if (off + c > 64 && limb < 3) v |= scalars[i][limb + 1] << (64 - off);
      v &= (u64(1) << c) - 1;
      if (v) buckets[v - 1] = buckets[v - 1].add_affine(bases[i]);
    }
    Jac<F> run = Jac<F>::identity();
    for (size_t j = buckets.size(); j-- > 0;) { run = run.add(buckets[j]); acc = acc.add(run); }
  }
  return acc;
}

template <class F>
Jac<F> best_multiexp(const u64 (*scalars)[4], const Affine<F>* bases, size_t n) {
  int nt = num_threads();
  if (nt <= 1 || n < (size_t)nt * 16) return multiexp_serial<F>(scalars, bases, n);
  size_t chunk = (n + nt - 1) / nt;
  std::vector<Jac<F>> parts((n + chunk - 1) / chunk);
  Pool::get().run(parts.size(), [&](size_t t) { size_t s = t * chunk, e = std::min(n, s + chunk); parts[t] = multiexp_serial<F>(scalars + s, bases + s, e - s); });
  Jac<F> acc = Jac<F>::identity();
  for (auto& p : parts) acc = acc.add(p);
  return acc;
}

// scalars in Montgomery form -> canonical limbs, then MSM
template <class F, class Sc>
Jac<F> msm(const Sc* scalars, const Affine<F>* bases, size_t n) {
  std::vector<u64[4]> can(n);
  parallel_for(n, [&](size_t s, size_t e) { for (size_t i = s; i < e; ++i) scalars[i].to_canonical(can[i]); });
  return best_multiexp<F>(can.data(), bases, n);
}

// ---- best_fft: in-place radix-2, natural order in and out: a[k] <- sum_i a[i] w^(ik)
template <class Sc>
void fft(Sc* a, int logn, const Sc& w) {
  size_t n = size_t(1) << logn;
  for (size_t i = 0; i < n; ++i) {
    size_t r = 0; for (int b = 0; b < logn; ++b) r |= ((i >> b) & 1) << (logn - 1 - b);
    if (i < r) std::swap(a[i], a[r]);
  }
  std::vector<Sc> tw(n / 2 ? n / 2 : 1);
  tw[0] = Sc::one();
  for (size_t i = 1; i < n / 2; ++i) tw[i] = tw[i - 1] * w;
  for (int s = 1; s <= logn; ++s) {
    size_t len = size_t(1) << s, half = len / 2, stride = n / len;
    parallel_for(n / 2, [&](size_t b0, size_t b1) {
      for (size_t b = b0; b < b1; ++b) {
        size_t blk = b / half, j = b % half;
        Sc* lo = a + blk * len + j; Sc* hi = lo + half;
        Sc t = *hi * tw[j * stride];
        *hi = *lo - t; *lo = *lo + t;
      }
    });
  }
}

template <class Sc>
struct Domain {  // halo2 EvaluationDomain::new(j = cs degree, k)
  int k, ext_k; size_t n, ext_n; uint32_t quotient_poly_degree;
  Sc omega, omega_inv, ext_omega, ext_omega_inv, zeta, zeta_inv, n_inv, ext_n_inv, barycentric_weight;
  std::vector<Sc> t_evaluations;  // inverses of t(zeta * ext_omega^i), period 2^(ext_k-k)
  Domain(uint32_t j, int k_) : k(k_) {
    quotient_poly_degree = j - 1;
    n = size_t(1) << k;
    ext_k = k; while ((size_t(1) << ext_k) < n * quotient_poly_degree) ++ext_k;
    ext_n = size_t(1) << ext_k;
    ext_omega = Sc::root_of_unity(); for (int i = ext_k; i < 32; ++i) ext_omega = ext_omega.sqr();
    omega = ext_omega; for (int i = k; i < ext_k; ++i) omega = omega.sqr();
    omega_inv = omega.inv(); ext_omega_inv = ext_omega.inv();
    zeta = Sc::zeta(); zeta_inv = zeta.sqr();
    n_inv = Sc::from_u64(n).inv(); ext_n_inv = Sc::from_u64(ext_n).inv();
    barycentric_weight = n_inv;
    Sc orig = zeta.pow_u64(n), step = ext_omega.pow_u64(n), cur = orig;
    do { t_evaluations.push_back(cur - Sc::one()); cur = cur * step; } while (cur != orig);
    batch_invert(t_evaluations.data(), t_evaluations.size());
  }
  void lagrange_to_coeff(std::vector<Sc>& a) const { fft(a.data(), k, omega_inv); for (auto& x : a) x = x * n_inv; }
  void coeff_to_lagrange(std::vector<Sc>& a) const { fft(a.data(), k, omega); }
  std::vector<Sc> coeff_to_extended(const std::vector<Sc>& c) const {
    std::vector<Sc> a(ext_n, Sc::zero());
    Sc zp[3] = {Sc::one(), zeta, zeta_inv};
    for (size_t i = 0; i < c.size(); ++i) a[i] = (i % 3) ? c[i] * zp[i % 3] : c[i];
    fft(a.data(), ext_k, ext_omega);
    return a;
  }
  std::vector<Sc> extended_to_coeff(std::vector<Sc> a) const {
    fft(a.data(), ext_k, ext_omega_inv);
    Sc zp[3] = {Sc::one(), zeta_inv, zeta};
    for (size_t i = 0; i < a.size(); ++i) { a[i] = a[i] * ext_n_inv; if (i % 3) a[i] = a[i] * zp[i % 3]; }
    a.resize(n * quotient_poly_degree);
    return a;
  }
  void divide_by_vanishing_poly(std::vector<Sc>& a) const {
    for (size_t i = 0; i < a.size(); ++i) a[i] = a[i] * t_evaluations[i % t_evaluations.size()];
  }
  Sc rotate_omega(const Sc& x, int rot) const {
    return rot >= 0 ? x * omega.pow_u64(rot) : x * omega_inv.pow_u64(-(int64_t)rot);
  }
};

template <class Sc> Sc eval_polynomial(const Sc* c, size_t n, const Sc& x) {
  int nt = num_threads();
  if (nt <= 1 || n < 1024) { Sc acc = Sc::zero(); for (size_t i = n; i-- > 0;) acc = acc * x + c[i]; return acc; }
  size_t chunk = (n + nt - 1) / nt, parts = (n + chunk - 1) / chunk;
  std::vector<Sc> res(parts);
  Pool::get().run(parts, [&](size_t t) {
    size_t s = t * chunk, e = std::min(n, s + chunk);
    Sc acc = Sc::zero(); for (size_t i = e; i-- > s;) acc = acc * x + c[i];
    res[t] = acc * x.pow_u64(s);
  });
  Sc acc = Sc::zero(); for (auto& r : res) acc = acc + r;
  return acc;
}

// halo2 kate_division: quotient of (a(X) - a(b)) / (X - b), length n-1
template <class Sc> std::vector<Sc> kate_division(const std::vector<Sc>& a, const Sc& b) {
  std::vector<Sc> q(a.size() - 1, Sc::zero());
  Sc tmp = Sc::zero();
  for (size_t i = a.size() - 1; i >= 1; --i) { Sc lead = a[i] + tmp * b; q[i - 1] = lead; tmp = lead; }
  return q;
}

}  // namespace orc
