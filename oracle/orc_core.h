// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path
// (finmath-lib_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may use it.
//
// CPU restatement (C++17, compiled with -ffp-contract=off so no FMA contraction, like the JVM) of the
// finmath-lib Monte-Carlo path-simulation hot path.  Citations: J/ = /root/reference/src/main/java/net/finmath/
//
// PARITY PINNING STATUS
//   * MT19937 words: pinned against the public MT19937 known-answer vector (init_by_array{0x123,0x234,0x345,0x456})
//     and against numpy's independent implementation (tests/golden/make_golden.py).  The arithmetic itself lives in
//     commons-math3 3.6.1 (R/pom.xml:22, jar only, no source in /root/reference): "parity unpinned" against the
//     reference itself, because no reference test asserts a specific random number and there is
//     no JVM in this image to run it.
//   * AS241: coefficients restated from Wichura (1988) as transcribed at J/functions/NormalDistribution.java:67-162,
//     checked against scipy.stats.norm.ppf; the reference has no golden values for it -> "parity unpinned".
//   * RandomVariable semantics: pinned by the reference's own exact-identity tests
//     (T/montecarlo/RandomVariableTest.java:53-142, restated in tests/test_oracle_rv.py).
//   * Model prices: pinned only by the reference tests' closed-form tolerances (see tests/test_oracle_models.py).
#pragma once
#include <cstdint>
#include <cmath>
#include <vector>
#include <memory>
#include <limits>
#include <algorithm>
#include <stdexcept>
#include <functional>

namespace orc {

// Math.exp / Math.log of the reference are libm calls (std::exp / std::log here, like the JVM's: < 1 ulp, not bit-specified).  Every
// exp / log of the restatement goes through these two pointers so that a DIAGNOSTIC mode can swap in another implementation
// (orc_set_math in orc_capi.cpp: the exp / log the device kernels use, compiled for the host).  With it the tests can attribute a
// deviation either to the arithmetic (none: the device then agrees to the last bit) or to the two libraries' last-bit differences.
typedef double (*MathFn)(double);
inline double stdExp(double x) { return std::exp(x); }
inline double stdLog(double x) { return std::log(x); }
inline MathFn& mathExp() { static MathFn f = stdExp; return f; }
inline MathFn& mathLog() { static MathFn f = stdLog; return f; }

// ---------------------------------------------------------------------------------------------------------------
// MT19937 as in org.apache.commons.math3.random.MersenneTwister 3.6.1 (bytecode-verified in SURVEY.md §8c) behind
// the wrapper J/randomnumbers/MersenneTwister.java:26-29 (seed ctor) and :48-51 (nextDoubleFast).
// ---------------------------------------------------------------------------------------------------------------
struct MersenneTwister {
	static constexpr int N = 624, M = 397;
	uint32_t mt[N];
	int mti;

	void initGenrand(uint32_t s) {                       // setSeed(int)
		mt[0] = s;
		for (mti = 1; mti < N; mti++) mt[mti] = 1812433253u * (mt[mti - 1] ^ (mt[mti - 1] >> 30)) + (uint32_t)mti;
	}
	void initByArray(const uint32_t* key, int len) {     // setSeed(int[])
		initGenrand(19650218u);
		int i = 1, j = 0;
		for (int k = std::max(N, len); k != 0; k--) {
			mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
			i++; j++;
			if (i >= N) { mt[0] = mt[N - 1]; i = 1; }
			if (j >= len) j = 0;
		}
		for (int k = N - 1; k != 0; k--) {
			mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
			i++;
			if (i >= N) { mt[0] = mt[N - 1]; i = 1; }
		}
		mt[0] = 0x80000000u;
		mti = N;
	}
	// new MersenneTwister(long seed): key = { (int)(seed >>> 32), (int)(seed & 0xffffffff) }
	explicit MersenneTwister(int64_t seed) {
		uint32_t key[2] = { (uint32_t)(((uint64_t)seed) >> 32), (uint32_t)(((uint64_t)seed) & 0xffffffffu) };
		initByArray(key, 2);
	}
	MersenneTwister(const uint32_t* key, int len) { initByArray(key, len); }

	void regenerate() {
		static const uint32_t mag01[2] = { 0u, 0x9908b0dfu };
		int k;
		for (k = 0; k < N - M; k++) {
			uint32_t y = (mt[k] & 0x80000000u) | (mt[k + 1] & 0x7fffffffu);
			mt[k] = mt[k + M] ^ (y >> 1) ^ mag01[y & 1u];
		}
		for (; k < N - 1; k++) {
			uint32_t y = (mt[k] & 0x80000000u) | (mt[k + 1] & 0x7fffffffu);
			mt[k] = mt[k + (M - N)] ^ (y >> 1) ^ mag01[y & 1u];
		}
		uint32_t y = (mt[N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
		mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ mag01[y & 1u];
		mti = 0;
	}
	uint32_t nextWord() {                                // next(32)
		if (mti >= N) regenerate();
		uint32_t y = mt[mti++];
		y ^= y >> 11;
		y ^= (y << 7) & 0x9d2c5680u;
		y ^= (y << 15) & 0xefc60000u;
		y ^= y >> 18;
		return y;
	}
	// BitsStreamGenerator.nextDouble(): ((long)next(26) << 26 | next(26)) * 0x1.0p-52
	double nextDouble() {
		uint64_t hi = (uint64_t)(nextWord() >> 6) << 26;
		uint64_t lo = (uint64_t)(nextWord() >> 6);
		return (double)(int64_t)(hi | lo) * 0x1.0p-52;
	}
	void skipWords(uint64_t n) {
		while (n > 0) {
			if (mti >= N) regenerate();
			uint64_t take = std::min<uint64_t>(n, (uint64_t)(N - mti));
			mti += (int)take; n -= take;
		}
	}
};

// ---------------------------------------------------------------------------------------------------------------
// AS241 PPND16, J/functions/NormalDistribution.java:67-162 (entry :47-49).  Expression order as there:
// q * num / den is (q*num)/den; Horner without contraction.  Quirk: r <= 0 returns 0.0 (:141-143).
// ---------------------------------------------------------------------------------------------------------------
inline double inverseCumulativeNormal(double p) {
	const double a0 = 3.3871328727963666080e+00, a1 = 1.3314166789178437745e+02, a2 = 1.9715909503065514427e+03,
		a3 = 1.3731693765509461125e+04, a4 = 4.5921953931549871457e+04, a5 = 6.7265770927008700853e+04,
		a6 = 3.3430575583588128105e+04, a7 = 2.5090809287301226727e+03;
	const double b1 = 4.2313330701600911252e+01, b2 = 6.8718700749205790830e+02, b3 = 5.3941960214247511077e+03,
		b4 = 2.1213794301586595867e+04, b5 = 3.9307895800092710610e+04, b6 = 2.8729085735721942674e+04,
		b7 = 5.2264952788528545610e+03;
	const double c0 = 1.42343711074968357734e+00, c1 = 4.63033784615654529590e+00, c2 = 5.76949722146069140550e+00,
		c3 = 3.64784832476320460504e+00, c4 = 1.27045825245236838258e+00, c5 = 2.41780725177450611770e-01,
		c6 = 2.27238449892691845833e-02, c7 = 7.74545014278341407640e-04;
	const double d1 = 2.05319162663775882187e+00, d2 = 1.67638483018380384940e+00, d3 = 6.89767334985100004550e-01,
		d4 = 1.48103976427480074590e-01, d5 = 1.51986665636164571966e-02, d6 = 5.47593808499534494600e-04,
		d7 = 1.05075007164441684324e-09;
	const double e0 = 6.65790464350110377720e+00, e1 = 5.46378491116411436990e+00, e2 = 1.78482653991729133580e+00,
		e3 = 2.96560571828504891230e-01, e4 = 2.65321895265761230930e-02, e5 = 1.24266094738807843860e-03,
		e6 = 2.71155556874348757815e-05, e7 = 2.01033439929228813265e-07;
	const double f1 = 5.99832206555887937690e-01, f2 = 1.36929880922735805310e-01, f3 = 1.48753612908506148525e-02,
		f4 = 7.86869131145613259100e-04, f5 = 1.84631831751005468180e-05, f6 = 1.42151175831644588870e-07,
		f7 = 2.04426310338993978564e-15;

	const double q = p - 0.5;
	double r;
	if (std::fabs(q) <= 0.425) {
		r = 0.180625 - q * q;
		return q * (((((((a7 * r + a6) * r + a5) * r + a4) * r + a3) * r + a2) * r + a1) * r + a0)
			/ (((((((b7 * r + b6) * r + b5) * r + b4) * r + b3) * r + b2) * r + b1) * r + 1.0);
	}
	r = (q < 0.0) ? p : 1.0 - p;
	if (r <= 0.0) return 0.0;
	r = std::sqrt(-mathLog()(r));
	double x;
	if (r <= 5.0) {
		r -= 1.6;
		x = (((((((c7 * r + c6) * r + c5) * r + c4) * r + c3) * r + c2) * r + c1) * r + c0)
			/ (((((((d7 * r + d6) * r + d5) * r + d4) * r + d3) * r + d2) * r + d1) * r + 1.0);
	} else {
		r -= 5.0;
		x = (((((((e7 * r + e6) * r + e5) * r + e4) * r + e3) * r + e2) * r + e1) * r + e0)
			/ (((((((f7 * r + f6) * r + f5) * r + f4) * r + f3) * r + f2) * r + f1) * r + 1.0);
	}
	return (q < 0.0) ? -x : x;
}

// ---------------------------------------------------------------------------------------------------------------
// TimeDiscretizationFromArray, J/time/TimeDiscretizationFromArray.java: tick rounding :387-389, distinct+sorted
// :57-64, equidistant ctor :210-217, getTimeStep :267-269, getTimeIndex (Arrays.binarySearch) :272-274.
// ---------------------------------------------------------------------------------------------------------------
struct TimeDiscretization {
	std::vector<double> t;
	double tick = 1.0 / (365.0 * 24.0);
	double roundTick(double x) const { return std::rint(x / tick) * tick; }
	TimeDiscretization() {}
	explicit TimeDiscretization(const std::vector<double>& times) {
		for (double x : times) t.push_back(roundTick(x));
		std::sort(t.begin(), t.end());
		t.erase(std::unique(t.begin(), t.end()), t.end());
	}
	TimeDiscretization(double initial, int numberOfTimeSteps, double deltaT) {
		std::vector<double> times;
		for (int n = 0; n <= numberOfTimeSteps; n++) times.push_back(initial + n * deltaT);
		*this = TimeDiscretization(times);
	}
	int getNumberOfTimes() const { return (int)t.size(); }
	int getNumberOfTimeSteps() const { return (int)t.size() - 1; }
	double getTime(int i) const { return t.at(i); }
	double getTimeStep(int i) const { return t.at(i + 1) - t.at(i); }
	// java.util.Arrays.binarySearch semantics: index if found, else -(insertionPoint)-1
	int getTimeIndex(double time) const {
		double key = roundTick(time);
		int lo = 0, hi = (int)t.size() - 1;
		while (lo <= hi) {
			int mid = (int)(((unsigned)lo + (unsigned)hi) >> 1);
			double v = t[mid];
			if (v < key) lo = mid + 1; else if (v > key) hi = mid - 1; else return mid;
		}
		return -(lo + 1);
	}
};

// ---------------------------------------------------------------------------------------------------------------
// RandomVariable semantics.  prio 0 = J/stochastic/Scalar.java (time -inf), prio 1 =
// J/montecarlo/RandomVariableFromDoubleArray.java.  Objects are immutable, shared by pointer.
// ---------------------------------------------------------------------------------------------------------------
struct RV;
using P = std::shared_ptr<const RV>;
static constexpr double NEG_INF = -std::numeric_limits<double>::infinity();

struct RV {
	int prio;
	double time;
	bool det;
	double v;
	std::vector<double> r;
	size_t size() const { return det ? 1 : r.size(); }
	double get(size_t i) const { return det ? v : r[i]; }
};

// counters so the CPU baseline can report "array passes" if wanted
inline P scalar(double v) { auto x = std::make_shared<RV>(); x->prio = 0; x->time = NEG_INF; x->det = true; x->v = v; return x; }
inline P rvconst(double time, double v) { auto x = std::make_shared<RV>(); x->prio = 1; x->time = time; x->det = true; x->v = v; return x; }
inline P rvvec(double time, std::vector<double>&& r) { auto x = std::make_shared<RV>(); x->prio = 1; x->time = time; x->det = false; x->v = std::numeric_limits<double>::quiet_NaN(); x->r = std::move(r); return x; }

// Java Math.min / Math.max (NaN-propagating, -0.0 < +0.0)
inline double jmin(double a, double b) {
	if (a != a) return a;
	if (a == 0.0 && b == 0.0 && std::signbit(b)) return b;
	return (a <= b) ? a : b;
}
inline double jmax(double a, double b) {
	if (a != a) return a;
	if (a == 0.0 && b == 0.0 && std::signbit(a)) return b;
	return (a >= b) ? a : b;
}

template <class F> inline P map1(const P& a, F f) {      // unary / rv∘double: keeps the receiver's time and type
	if (a->det) return a->prio == 0 ? scalar(f(a->v)) : rvconst(a->time, f(a->v));
	std::vector<double> o(a->r.size());
	for (size_t i = 0; i < o.size(); i++) o[i] = f(a->r[i]);
	return rvvec(a->time, std::move(o));
}
template <class F> inline P vec2(double time, const P& a, const P& b, F f) {
	size_t n = std::max(a->size(), b->size());
	std::vector<double> o(n);
	for (size_t i = 0; i < n; i++) o[i] = f(a->get(i), b->get(i));
	return rvvec(time, std::move(o));
}
template <class F> inline P vec3(double time, const P& a, const P& b, const P& c, F f) {
	size_t n = std::max(std::max(a->size(), b->size()), c->size());
	std::vector<double> o(n);
	for (size_t i = 0; i < n; i++) o[i] = f(a->get(i), b->get(i), c->get(i));
	return rvvec(time, std::move(o));
}

// rv ∘ double  (RandomVariableFromDoubleArray.java:742-875, Scalar.java same formulas)
inline P add(const P& a, double x) { return map1(a, [x](double y) { return y + x; }); }
inline P sub(const P& a, double x) { return map1(a, [x](double y) { return y - x; }); }
inline P bus(const P& a, double x) { return map1(a, [x](double y) { return x - y; }); }
inline P mult(const P& a, double x) { return map1(a, [x](double y) { return y * x; }); }
inline P div(const P& a, double x) { return map1(a, [x](double y) { return y / x; }); }
inline P vid(const P& a, double x) { return map1(a, [x](double y) { return x / y; }); }
inline P cap(const P& a, double x) { return map1(a, [x](double y) { return jmin(y, x); }); }
inline P floor(const P& a, double x) { return map1(a, [x](double y) { return jmax(y, x); }); }
// Math.pow: the reference's own tests require pow(x, 2.0) == x * x and pow(x, 0.5) == sqrt(x) bit for bit
// (T/montecarlo/RandomVariableTest.java:101-127; HotSpot's pow intrinsic special-cases both exponents), which a libm pow does not
// guarantee in the last bit - restated explicitly.
inline P pow(const P& a, double x) {
	if (x == 2.0) return map1(a, [](double y) { return y * y; });
	if (x == 0.5) return map1(a, [](double y) { return std::sqrt(y); });
	return map1(a, [x](double y) { return std::pow(y, x); });
}
inline P squared(const P& a) { return map1(a, [](double y) { return y * y; }); }
inline P sqrt(const P& a) { return map1(a, [](double y) { return std::sqrt(y); }); }
inline P exp(const P& a) { return map1(a, [](double y) { return mathExp()(y); }); }
inline P expm1(const P& a) { return map1(a, [](double y) { return std::expm1(y); }); }
inline P log(const P& a) { return map1(a, [](double y) { return mathLog()(y); }); }
inline P sin(const P& a) { return map1(a, [](double y) { return std::sin(y); }); }
inline P cos(const P& a) { return map1(a, [](double y) { return std::cos(y); }); }
inline P invert(const P& a) { return map1(a, [](double y) { return 1.0 / y; }); }
inline P abs(const P& a) { return map1(a, [](double y) { return std::fabs(y); }); }
inline P isNaN(const P& a) { return map1(a, [](double y) { return (y != y) ? 1.0 : 0.0; }); }

inline double tmax(const P& a, const P& b) { return std::max(a->time, b->time); }

// rv ∘ rv — Scalar.java:276-312 for a Scalar receiver, RandomVariableFromDoubleArray.java:1027-1276 otherwise
inline P add(const P& a, const P& b) {
	if (a->prio == 0) return add(b, a->v);
	if (b->prio > a->prio) return add(b, a);
	if (a->det && b->det) return rvconst(tmax(a, b), a->v + b->v);
	if (a->det) return vec2(tmax(a, b), a, b, [](double x, double y) { return x + y; });
	if (b->det) return add(a, b->v);
	return vec2(tmax(a, b), a, b, [](double x, double y) { return x + y; });
}
inline P sub(const P& a, const P& b) {
	if (a->prio == 0) return mult(sub(b, a->v), -1.0);
	if (a->det && b->det) return rvconst(tmax(a, b), a->v - b->v);
	if (a->det) return vec2(tmax(a, b), a, b, [](double x, double y) { return x - y; });
	if (b->det) return sub(a, b->v);
	return vec2(tmax(a, b), a, b, [](double x, double y) { return x - y; });
}
inline P bus(const P& a, const P& b) {                   // b - a
	if (a->prio == 0) return sub(b, a->v);
	if (a->det && b->det) return rvconst(tmax(a, b), b->v - a->v);
	return vec2(tmax(a, b), a, b, [](double x, double y) { return y - x; });
}
inline P mult(const P& a, const P& b) {
	if (a->prio == 0) return mult(b, a->v);
	if (a->det && b->det) return rvconst(tmax(a, b), a->v * b->v);
	if (b->det) return mult(a, b->v);
	return vec2(tmax(a, b), a, b, [](double x, double y) { return x * y; });
}
inline P div(const P& a, const P& b) {
	if (a->prio == 0) return mult(invert(b), a->v);
	if (a->det && b->det) return rvconst(tmax(a, b), a->v / b->v);
	return vec2(tmax(a, b), a, b, [](double x, double y) { return x / y; });
}
inline P vid(const P& a, const P& b) {                   // b / a
	if (a->prio == 0) return div(b, a->v);
	if (a->det && b->det) return rvconst(tmax(a, b), b->v / a->v);
	return vec2(tmax(a, b), a, b, [](double x, double y) { return y / x; });
}
inline P cap(const P& a, const P& b) {
	if (a->prio == 0) return cap(b, a->v);
	if (a->det && b->det) return rvconst(tmax(a, b), jmin(a->v, b->v));
	return vec2(tmax(a, b), a, b, [](double x, double y) { return jmin(x, y); });
}
inline P floor(const P& a, const P& b) {
	if (a->prio == 0) return floor(b, a->v);
	if (a->det && b->det) return rvconst(tmax(a, b), jmax(a->v, b->v));
	if (!a->det && b->det) return floor(a, b->v);
	return vec2(tmax(a, b), a, b, [](double x, double y) { return jmax(x, y); });
}
// Scalar.java:315-328, RandomVariableFromDoubleArray.java:1278-1334
inline P accrue(const P& a, const P& rate, double pl) {
	if (a->prio == 0) return add(mult(rate, pl * a->v), a->v);
	if (rate->det) return mult(a, 1.0 + rate->v * pl);
	return vec2(tmax(a, rate), a, rate, [pl](double x, double r) { return x * (1 + r * pl); });
}
inline P discount(const P& a, const P& rate, double pl) {
	if (a->prio == 0) {
		if (a->v == 0) return mult(rate, 0.0);
		return invert(add(mult(rate, pl / a->v), 1.0 / a->v));
	}
	if (rate->det) return div(a, 1.0 + rate->v * pl);
	return vec2(tmax(a, rate), a, rate, [pl](double x, double r) { return x / (1.0 + r * pl); });
}
// Scalar.java:330-337, RandomVariableFromDoubleArray.java:1341-1363
inline P choose(const P& trigger, const P& nonNeg, const P& neg) {
	if (trigger->det) return (trigger->v >= 0) ? nonNeg : neg;
	double t = std::max(std::max(trigger->time, nonNeg->time), neg->time);
	return vec3(t, trigger, nonNeg, neg, [](double c, double x, double y) { return c >= 0.0 ? x : y; });
}
// Scalar.java:349-357, RandomVariableFromDoubleArray.java:1365-1427
inline P addProduct(const P& a, const P& f1, double f2) {
	if (a->prio == 0) return add(mult(f1, f2), a->v);
	if (f1->det) return add(a, f1->v * f2);
	return vec2(tmax(a, f1), a, f1, [f2](double x, double y) { return x + y * f2; });
}
inline P addProduct(const P& a, const P& f1, const P& f2) {
	if (a->prio == 0) return add(mult(f1, f2), a->v);
	double t = std::max(std::max(a->time, f1->time), f2->time);
	if (a->det && f1->det && f2->det) return rvconst(t, a->v + (f1->v * f2->v));
	if (f1->det && f2->det) return add(a, f1->v * f2->v);
	if (f2->det) return addProduct(a, f1, f2->v);
	if (f1->det) return addProduct(a, f2, f1->v);
	if (!a->det) return vec3(t, a, f1, f2, [](double x, double y, double z) { return x + y * z; });
	return add(a, mult(f1, f2));
}
inline P addRatio(const P& a, const P& num, const P& den) {
	if (a->prio == 0) return add(div(num, den), a->v);
	double t = std::max(std::max(a->time, num->time), den->time);
	if (a->det && num->det && den->det) return rvconst(t, a->v + (num->v / den->v));
	return vec3(t, a, num, den, [](double x, double y, double z) { return x + y / z; });
}
inline P subRatio(const P& a, const P& num, const P& den) {
	if (a->prio == 0) return mult(sub(div(num, den), a->v), -1.0);
	double t = std::max(std::max(a->time, num->time), den->time);
	if (a->det && num->det && den->det) return rvconst(t, a->v - (num->v / den->v));
	return vec3(t, a, num, den, [](double x, double y, double z) { return x - y / z; });
}
// RandomVariable.java:659-666 (default method)
inline P addSumProduct(const P& a, const std::vector<P>& f1, const std::vector<P>& f2) {
	P result = a;
	for (size_t i = 0; i < f1.size(); i++) result = addProduct(result, f1[i], f2[i]);
	return result;
}

// Reductions — RandomVariableFromDoubleArray.java:286-428 (sequential Kahan), Scalar.java:100-135
inline double kahanMean(const std::vector<double>& r) {
	double sum = 0.0, error = 0.0;
	for (size_t i = 0; i < r.size(); i++) {
		const double value = r[i] - error;
		const double newSum = sum + value;
		error = (newSum - sum) - value;
		sum = newSum;
	}
	return sum / (double)r.size();
}
inline double getAverage(const P& a) {
	if (a->det) return a->v;
	if (a->r.empty()) return std::numeric_limits<double>::quiet_NaN();
	return kahanMean(a->r);
}
inline double getAverage(const P& a, const P& prob) {
	if (a->det) return a->v * getAverage(prob);
	if (a->r.empty()) return std::numeric_limits<double>::quiet_NaN();
	double sum = 0.0, error = 0.0;
	for (size_t i = 0; i < a->r.size(); i++) {
		const double value = a->r[i] * prob->get(i) - error;
		const double newSum = sum + value;
		error = (newSum - sum) - value;
		sum = newSum;
	}
	return sum / (double)a->r.size();
}
inline double getVariance(const P& a) {
	if (a->det || a->size() == 1) return 0.0;
	if (a->r.empty()) return std::numeric_limits<double>::quiet_NaN();
	const double average = getAverage(a);
	double sum = 0.0, err = 0.0;
	for (size_t i = 0; i < a->r.size(); i++) {
		const double value = (a->r[i] - average) * (a->r[i] - average) - err;
		const double newSum = sum + value;
		err = (newSum - sum) - value;
		sum = newSum;
	}
	return sum / (double)a->r.size();
}
inline double getVariance(const P& a, const P& prob) {   // quirk: NOT divided by n (:379)
	if (a->det) return 0.0;
	if (a->r.empty()) return std::numeric_limits<double>::quiet_NaN();
	const double average = getAverage(a, prob);
	double sum = 0.0, err = 0.0;
	for (size_t i = 0; i < a->r.size(); i++) {
		const double value = (a->r[i] - average) * (a->r[i] - average) * prob->get(i) - err;
		const double newSum = sum + value;
		err = (newSum - sum) - value;
		sum = newSum;
	}
	return sum;
}
inline double getSampleVariance(const P& a) {
	if (a->det || a->size() == 1) return 0.0;
	if (a->r.empty()) return std::numeric_limits<double>::quiet_NaN();
	return getVariance(a) * (double)a->size() / (double)(a->size() - 1);
}
inline double getStandardDeviation(const P& a) {
	if (a->det) return 0.0;
	if (a->r.empty()) return std::numeric_limits<double>::quiet_NaN();
	return std::sqrt(getVariance(a));
}
inline double getStandardError(const P& a) {
	if (a->det) return 0.0;
	if (a->r.empty()) return std::numeric_limits<double>::quiet_NaN();
	return getStandardDeviation(a) / std::sqrt((double)a->size());
}
inline double getMin(const P& a) {
	if (a->det) return a->v;
	double m = std::numeric_limits<double>::max();
	if (!a->r.empty()) m = a->r[0];
	for (double x : a->r) m = jmin(x, m);
	return m;
}
inline double getMax(const P& a) {
	if (a->det) return a->v;
	double m = -std::numeric_limits<double>::max();
	if (!a->r.empty()) m = a->r[0];
	for (double x : a->r) m = jmax(x, m);
	return m;
}
inline long jround(double x) { return (long)std::floor(x + 0.5); }   // Math.round
inline double getQuantile(const P& a, double q) {        // :445-460
	if (a->det) return a->v;
	if (a->r.empty()) return std::numeric_limits<double>::quiet_NaN();
	std::vector<double> s = a->r;
	std::sort(s.begin(), s.end());
	long n = (long)s.size();
	long idx = std::min(std::max((long)(int)jround((n + 1) * q - 1), 0L), n - 1);
	return s[idx];
}
inline double getQuantileExpectation(const P& a, double qs, double qe) {   // :474-499
	if (a->det) return a->v;
	if (a->r.empty()) return std::numeric_limits<double>::quiet_NaN();
	if (qs > qe) return getQuantileExpectation(a, qe, qs);
	std::vector<double> s = a->r;
	std::sort(s.begin(), s.end());
	long n = (long)s.size();
	long i0 = std::min(std::max((long)(int)jround((n + 1) * qs - 1), 0L), n - 1);
	long i1 = std::min(std::max((long)(int)jround((n + 1) * qe - 1), 0L), n - 1);
	double e = 0.0;
	for (long i = i0; i <= i1; i++) e += s[i];
	return e / (double)(i1 - i0 + 1);
}
inline std::vector<double> getHistogram(const P& a, const std::vector<double>& pts) {   // :501-550
	std::vector<double> h(pts.size() + 1, 0.0);
	if (a->det) {
		for (size_t k = 0; k < pts.size(); k++) if (a->v > pts[k]) { h[k] = 1.0; break; }
		h[pts.size()] = 1.0;
		return h;
	}
	std::vector<double> s = a->r;
	std::sort(s.begin(), s.end());
	size_t idx = 0;
	for (size_t k = 0; k < pts.size(); k++) {
		int c = 0;
		while (idx < s.size() && s[idx] <= pts[k]) { idx++; c++; }
		h[k] = c;
	}
	h[pts.size()] = (double)(s.size() - idx);
	if (!s.empty()) for (double& x : h) x /= (double)s.size();
	return h;
}

} // namespace orc
