// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_core.h header).  extern "C" surface used by oracle/oracle.py (ctypes).
// Build: make -C oracle   (g++ -O2 -ffp-contract=off; no CUDA, no reference sources).
#include "orc_products.h"
#include <cstring>
#include <thread>
#include <chrono>

using namespace orc;

// Diagnostic only: the device kernels' exp / log (finmath-lib_b200/csrc/fmb_math.cuh, host-compilable) for orc_set_math(1).
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wunknown-pragmas"
#include "../finmath-lib_b200/csrc/fmb_math.cuh"
#pragma GCC diagnostic pop
static double deviceExp(double x) { return fmb::fexp(x); }
static double deviceLog(double x) { return fmb::flog(x); }

namespace {
std::vector<double> vecOf(const double* p, int n) { return std::vector<double>(p, p + n); }
P wrap(const double* x, uint64_t n) { return rvvec(0.0, std::vector<double>(x, x + n)); }
void store(const P& r, double* out, uint64_t n) {
	for (uint64_t i = 0; i < n; i++) out[i] = r->get(i);
}
}

extern "C" {

// ---- MT19937 / AS241 / time grid ---------------------------------------------------------------------------
void orc_mt_words(int64_t seed, uint64_t offset, uint64_t n, uint32_t* out) {
	MersenneTwister mt(seed);
	mt.skipWords(offset);
	for (uint64_t i = 0; i < n; i++) out[i] = mt.nextWord();
}
void orc_mt_words_key(const uint32_t* key, int len, uint64_t n, uint32_t* out) {
	MersenneTwister mt(key, len);
	for (uint64_t i = 0; i < n; i++) out[i] = mt.nextWord();
}
// raw (untempered) state words x[k], k >= 0, of the linear recurrence (x[0..623] = seeded state)
void orc_mt_raw_sequence(int64_t seed, uint64_t n, uint32_t* out) {
	MersenneTwister mt(seed);
	uint64_t k = 0;
	for (; k < 624 && k < n; k++) out[k] = mt.mt[k];
	while (k < n) {
		mt.regenerate();
		for (int i = 0; i < 624 && k < n; i++, k++) out[k] = mt.mt[i];
	}
}
void orc_mt_uniforms(int64_t seed, uint64_t offsetUniforms, uint64_t n, double* out) {
	MersenneTwister mt(seed);
	mt.skipWords(2 * offsetUniforms);
	for (uint64_t i = 0; i < n; i++) out[i] = mt.nextDouble();
}
void orc_icdf(const double* p, uint64_t n, double* out) {
	for (uint64_t i = 0; i < n; i++) out[i] = inverseCumulativeNormal(p[i]);
}
int orc_time_discretization(double initial, int nSteps, double dt, double* out) {
	TimeDiscretization td(initial, nSteps, dt);
	for (size_t i = 0; i < td.t.size(); i++) out[i] = td.t[i];
	return (int)td.t.size();
}
int orc_time_discretization_from_array(const double* in, int n, double* out) {
	TimeDiscretization td(vecOf(in, n));
	for (size_t i = 0; i < td.t.size(); i++) out[i] = td.t[i];
	return (int)td.t.size();
}
int orc_time_index(const double* times, int n, double t) {
	TimeDiscretization td; td.t = vecOf(times, n);
	return td.getTimeIndex(t);
}
static TimeDiscretization tdFrom(const double* times, int nTimes) { TimeDiscretization td; td.t = vecOf(times, nTimes); return td; }

// out[t][f][p]
void orc_brownian(int seed, const double* times, int nTimes, int F, int paths, int64_t pathOffset, double* out) {
	BrownianMotion bm(tdFrom(times, nTimes), F, paths, seed, pathOffset);
	const int T = nTimes - 1;
	for (int t = 0; t < T; t++) for (int f = 0; f < F; f++)
		std::memcpy(out + ((size_t)t * F + f) * paths, bm.inc[t][f]->r.data(), sizeof(double) * paths);
}

// ---- element-wise semantics (op codes = include/finmath_b200.h) -----------------------------------------------
int orc_rv_unary(int op, const double* x, uint64_t n, double a, double* out) {
	P X = wrap(x, n), R;
	switch (op) {
	case 0: R = squared(X); break; case 1: R = sqrt(X); break; case 2: R = exp(X); break; case 3: R = log(X); break;
	case 4: R = sin(X); break; case 5: R = cos(X); break; case 6: R = invert(X); break; case 7: R = abs(X); break;
	case 8: R = isNaN(X); break; case 9: R = expm1(X); break;
	case 10: R = add(X, a); break; case 11: R = sub(X, a); break; case 12: R = bus(X, a); break; case 13: R = mult(X, a); break;
	case 14: R = div(X, a); break; case 15: R = vid(X, a); break; case 16: R = cap(X, a); break; case 17: R = floor(X, a); break;
	case 18: R = pow(X, a); break;
	default: return 1;
	}
	store(R, out, n); return 0;
}
int orc_rv_binary(int op, const double* x, const double* y, uint64_t n, double* out) {
	P X = wrap(x, n), Y = wrap(y, n), R;
	switch (op) {
	case 0: R = add(X, Y); break; case 1: R = sub(X, Y); break; case 2: R = mult(X, Y); break; case 3: R = div(X, Y); break;
	case 4: R = cap(X, Y); break; case 5: R = floor(X, Y); break;
	default: return 1;
	}
	store(R, out, n); return 0;
}
int orc_rv_ternary(int op, const double* x, const double* y, const double* z, uint64_t n, double a, double* out) {
	P X = wrap(x, n), Y = wrap(y, n), Z = z ? wrap(z, n) : P(), R;
	switch (op) {
	case 0: R = addProduct(X, Y, Z); break; case 1: R = addProduct(X, Y, a); break; case 2: R = addRatio(X, Y, Z); break;
	case 3: R = subRatio(X, Y, Z); break; case 4: R = accrue(X, Y, a); break; case 5: R = discount(X, Y, a); break;
	case 6: R = choose(X, Y, Z); break;
	default: return 1;
	}
	store(R, out, n); return 0;
}
// reductions in the reference's own (sequential Kahan) arithmetic
double orc_rv_reduce(int op, const double* x, const double* w, uint64_t n, double a, double b) {
	P X = wrap(x, n), W = w ? wrap(w, n) : P();
	switch (op) {
	case 0: return getAverage(X);
	case 1: return getAverage(X, W);
	case 2: return getVariance(X);
	case 3: return getVariance(X, W);
	case 4: return getMin(X);
	case 5: return getMax(X);
	case 6: return getSampleVariance(X);
	case 7: return getStandardDeviation(X);
	case 8: return getStandardError(X);
	case 9: return getQuantile(X, a);
	case 10: return getQuantileExpectation(X, a, b);
	}
	return std::numeric_limits<double>::quiet_NaN();
}
void orc_rv_histogram(const double* x, uint64_t n, const double* pts, int npts, double* out) {
	std::vector<double> h = getHistogram(wrap(x, n), vecOf(pts, npts));
	for (size_t i = 0; i < h.size(); i++) out[i] = h[i];
}
void orc_solve_pinv(const double* A, const double* b, int K, double* x, double* cond) {
	std::vector<double> r = solveSymmetricPseudoInverse(vecOf(A, K * K), vecOf(b, K), K, cond);
	for (int i = 0; i < K; i++) x[i] = r[i];
}

// ---- Black-Scholes / Heston ------------------------------------------------------------------------------
static void dumpProcess(Process& pr, int T, int N, int paths, double* out) {
	for (int t = 0; t <= T; t++) for (int c = 0; c < N; c++) {
		P x = pr.getProcessValue(t, c);
		double* o = out + ((size_t)t * N + c) * paths;
		for (int p = 0; p < paths; p++) o[p] = x->get(p);
	}
}
// out[(T+1)][1][P] (may be NULL); returns the EuropeanOption value average (J/.../EuropeanOption.java:172-193)
double orc_bs_european(int seed, const double* times, int nTimes, int paths, int64_t pathOffset, double s0, double r, double sigma,
		int scheme, double maturity, double strike, int callPut, double* processOut, double* valuesOut) {
	BrownianMotion bm(tdFrom(times, nTimes), 1, paths, seed, pathOffset);
	BlackScholesModel m(s0, r, sigma);
	Process pr(&m, &bm, scheme);
	if (processOut) dumpProcess(pr, nTimes - 1, 1, paths, processOut);
	P v = europeanOptionValue(m, pr, 0.0, maturity, strike, callPut);
	if (valuesOut) store(v, valuesOut, paths);
	return getAverage(v);
}
double orc_heston_european(int seed, const double* times, int nTimes, int paths, int64_t pathOffset, double s0, double r, double sigma,
		double discountRate, double theta, double kappa, double xi, double rho, int hestonScheme, int scheme,
		double maturity, double strike, int callPut, double* processOut, double* valuesOut) {
	BrownianMotion bm(tdFrom(times, nTimes), 2, paths, seed, pathOffset);
	HestonModel m(s0, r, sigma, discountRate, theta, kappa, xi, rho, hestonScheme);
	Process pr(&m, &bm, scheme);
	if (processOut) dumpProcess(pr, nTimes - 1, 2, paths, processOut);
	P v = europeanOptionValue(m, pr, 0.0, maturity, strike, callPut);
	if (valuesOut) store(v, valuesOut, paths);
	return getAverage(v);
}

// ---- LIBOR market model (handle based) -------------------------------------------------------------------
struct LmmHandle {
	std::unique_ptr<BrownianMotion> bm;
	LIBORMarketModel model;
	std::unique_ptr<Process> process;
	LIBORSimulation sim;
	BermudanResult lastBermudan;
};
void* orc_lmm_create(int seed, const double* simTimes, int nSimTimes, const double* tenorTimes, int nTenorTimes, int F, int paths,
		int64_t pathOffset, const double* L0, const double* discountFactors /* nTenorTimes or NULL */,
		const double* sigma /* [T][N] */, const double* factorMatrix /* [N][F] */, int measure, int stateSpace, double liborCap, int scheme) {
	auto* h = new LmmHandle();
	h->bm.reset(new BrownianMotion(tdFrom(simTimes, nSimTimes), F, paths, seed, pathOffset));
	const int N = nTenorTimes - 1, T = nSimTimes - 1;
	h->model.tenor = tdFrom(tenorTimes, nTenorTimes);
	h->model.L0 = vecOf(L0, N);
	if (discountFactors) h->model.discountFactors = vecOf(discountFactors, nTenorTimes);
	h->model.sigma = vecOf(sigma, T * N);
	h->model.factorMatrix = vecOf(factorMatrix, N * F);
	h->model.F = F; h->model.measure = measure; h->model.stateSpace = stateSpace; h->model.liborCap = liborCap;
	h->process.reset(new Process(&h->model, h->bm.get(), scheme));
	h->sim.model = &h->model; h->sim.process = h->process.get();
	return h;
}
void orc_lmm_free(void* hv) { delete (LmmHandle*)hv; }
void orc_lmm_process(void* hv, double* out /* [T+1][N][P] */) {
	auto* h = (LmmHandle*)hv;
	dumpProcess(*h->process, h->bm->td.getNumberOfTimeSteps(), h->model.getNumberOfComponents(), h->bm->paths, out);
}
void orc_lmm_brownian(void* hv, double* out /* [T][F][P] */) {
	auto* h = (LmmHandle*)hv;
	const int T = h->bm->td.getNumberOfTimeSteps();
	for (int t = 0; t < T; t++) for (int f = 0; f < h->bm->F; f++)
		std::memcpy(out + ((size_t)t * h->bm->F + f) * h->bm->paths, h->bm->inc[t][f]->r.data(), sizeof(double) * h->bm->paths);
}
// 0: libm exp / log (the default, what "the oracle" means everywhere).  1: the device's exp / log restated on the host - used by the
// tests to show that what separates the device paths from the oracle's is ONLY the last-bit difference of two exp / log libraries.
void orc_set_math(int mode) {
	mathExp() = mode == 1 ? deviceExp : stdExp;
	mathLog() = mode == 1 ? deviceLog : stdLog;
}
void orc_lmm_set_interpolation(void* hv, int method) { ((LmmHandle*)hv)->model.interpolationMethod = method; }
void orc_lmm_numeraire(void* hv, double time, double* out) {
	auto* h = (LmmHandle*)hv;
	store(h->sim.getNumeraire(time), out, h->bm->paths);
}
void orc_lmm_forward_rate(void* hv, double time, double start, double end, double* out) {
	auto* h = (LmmHandle*)hv;
	store(h->sim.getForwardRate(time, start, end), out, h->bm->paths);
}
double orc_lmm_swaption(void* hv, double exerciseDate, const double* fixingDates, const double* paymentDates, const double* swaprates,
		int n, double notional, double* valuesOut, double* stdErrOut) {
	auto* h = (LmmHandle*)hv;
	P v = swaptionValue(h->sim, 0.0, exerciseDate, vecOf(fixingDates, n), vecOf(paymentDates, n), vecOf(swaprates, n), notional, {});
	if (valuesOut) store(v, valuesOut, h->bm->paths);
	if (stdErrOut) *stdErrOut = getStandardError(v);
	return getAverage(v);
}
// the same with per-period discounting adjustments forwardBondOnForwardCurve / forwardBondOnDiscountCurve (Swaption.java:160-171)
double orc_lmm_swaption_adj(void* hv, double exerciseDate, const double* fixingDates, const double* paymentDates, const double* swaprates,
		int n, double notional, const double* adjustments, double* valuesOut, double* stdErrOut) {
	auto* h = (LmmHandle*)hv;
	P v = swaptionValue(h->sim, 0.0, exerciseDate, vecOf(fixingDates, n), vecOf(paymentDates, n), vecOf(swaprates, n), notional, vecOf(adjustments, n));
	if (valuesOut) store(v, valuesOut, h->bm->paths);
	if (stdErrOut) *stdErrOut = getStandardError(v);
	return getAverage(v);
}
double orc_lmm_caplet(void* hv, double maturity, double periodLength, double strike, double daycountFraction, int isFloorlet, double* valuesOut) {
	auto* h = (LmmHandle*)hv;
	P v = capletValue(h->sim, 0.0, maturity, periodLength, strike, daycountFraction, isFloorlet != 0);
	if (valuesOut) store(v, valuesOut, h->bm->paths);
	return getAverage(v);
}
// returns price; regressionOut[nExercise][6] in backward-loop order; condOut[nExercise]; exerciseTimeOut[P]
double orc_lmm_bermudan(void* hv, const int* isExercise, const double* fixingDates, const double* periodLengths, const double* paymentDates,
		const double* notionals, const double* swaprates, int n, int isCallable, double* valuesOut, double* exerciseTimeOut,
		double* regressionOut, double* condOut, double* stdErrOut) {
	auto* h = (LmmHandle*)hv;
	BermudanResult r = bermudanSwaptionValues(h->sim, 0.0, std::vector<int>(isExercise, isExercise + n), vecOf(fixingDates, n),
		vecOf(periodLengths, n), vecOf(paymentDates, n), vecOf(notionals, n), vecOf(swaprates, n), isCallable != 0);
	if (valuesOut) store(r.value, valuesOut, h->bm->paths);
	if (exerciseTimeOut) store(r.exerciseTime, exerciseTimeOut, h->bm->paths);
	if (regressionOut) for (size_t e = 0; e < r.regressionParameters.size(); e++)
		for (size_t k = 0; k < r.regressionParameters[e].size(); k++) regressionOut[e * 6 + k] = r.regressionParameters[e][k];
	if (condOut) for (size_t e = 0; e < r.regressionCond.size(); e++) condOut[e] = r.regressionCond[e];
	if (stdErrOut) *stdErrOut = getStandardError(r.value);
	return getAverage(r.value);
}

// replay with given regression coefficients ([nExercise][6], loop order) and Monte-Carlo weight: per-path values / exercise times of a window
void orc_lmm_bermudan_given(void* hv, const int* isExercise, const double* fixingDates, const double* periodLengths, const double* paymentDates,
		const double* notionals, const double* swaprates, int n, int isCallable, const double* coefficients, int nExercise, double weight,
		double* valuesOut, double* exerciseTimeOut) {
	auto* h = (LmmHandle*)hv;
	std::vector<std::vector<double>> coef(nExercise);
	for (int e = 0; e < nExercise; e++) coef[e] = vecOf(coefficients + 6 * e, 6);
	BermudanResult r = bermudanSwaptionValues(h->sim, 0.0, std::vector<int>(isExercise, isExercise + n), vecOf(fixingDates, n),
		vecOf(periodLengths, n), vecOf(paymentDates, n), vecOf(notionals, n), vecOf(swaprates, n), isCallable != 0, &coef, weight);
	if (valuesOut) store(r.value, valuesOut, h->bm->paths);
	if (exerciseTimeOut) store(r.exerciseTime, exerciseTimeOut, h->bm->paths);
}
// the six regression basis functions of BermudanSwaption.getBasisFunctions (:215-252) at one exercise date: out[6][P]
void orc_lmm_bermudan_basis(void* hv, double fixingDate, const double* fixingDates, const double* paymentDates, int n, double* out) {
	auto* h = (LmmHandle*)hv;
	std::vector<P> b = bermudanBasisFunctions(h->sim, fixingDate, vecOf(fixingDates, n), vecOf(paymentDates, n));
	for (size_t k = 0; k < b.size(); k++) store(b[k], out + k * (size_t)h->bm->paths, h->bm->paths);
}

// ---- Hull-White (process values only; [T+1][2][P]) -----------------------------------------------------------------
void orc_hull_white_process(int seed, const double* times, int nTimes, int paths, int64_t pathOffset, const double* volTimes, int nVolTimes,
		const double* vol, const double* mr, int scheme, double* processOut, double* coefOut /* [T][6]: c0, c1, l00, l01, l10, l11 or NULL */) {
	BrownianMotion bm(tdFrom(times, nTimes), 2, paths, seed, pathOffset);
	HullWhiteModel m;
	m.volTimes = tdFrom(volTimes, nVolTimes);
	m.vol = vecOf(vol, nVolTimes); m.mr = vecOf(mr, nVolTimes);
	Process pr(&m, &bm, scheme);
	dumpProcess(pr, nTimes - 1, 2, paths, processOut);
	if (coefOut) {
		std::vector<P> one = { scalar(1.0), scalar(0.0) };
		for (int t = 0; t < nTimes - 1; t++) {
			std::vector<P> d = m.getDrift(pr, t, one);
			std::vector<P> f0 = m.getFactorLoading(pr, t, 0, one), f1 = m.getFactorLoading(pr, t, 1, one);
			double* c = coefOut + 6 * t;
			c[0] = d[0]->v; c[1] = d[1]->v; c[2] = f0[0]->v; c[3] = f0[1]->v; c[4] = f1[0]->v; c[5] = f1[1]->v;
		}
	}
}

// Hull-White caplet (Caplet.java:114-160 on the Hull-White model's getForwardRate / getNumeraire); returns the price, fills values
double orc_hull_white_caplet(int seed, const double* times, int nTimes, int paths, const double* volTimes, int nVolTimes, const double* vol, const double* mr,
		const double* curveTimes, int nCurveTimes, const double* dfDiscount /* or NULL */, const double* dfForward, int scheme,
		double maturity, double periodLength, double strike, double* valuesOut, double* numeraireOut /* at maturity+periodLength */, double* forwardRateOut) {
	BrownianMotion bm(tdFrom(times, nTimes), 2, paths, seed, 0);
	HullWhiteModel m;
	m.volTimes = tdFrom(volTimes, nVolTimes);
	m.vol = vecOf(vol, nVolTimes); m.mr = vecOf(mr, nVolTimes);
	m.curveTimes = tdFrom(curveTimes, nCurveTimes);
	if (dfDiscount) m.dfDiscount = vecOf(dfDiscount, nCurveTimes);
	m.dfForward = vecOf(dfForward, nCurveTimes);
	Process pr(&m, &bm, scheme);
	const double paymentDate = maturity + periodLength;
	P fr = m.getForwardRate(pr, maturity, maturity, paymentDate);
	P numeraire = m.getNumeraire(pr, paymentDate);
	P w = pr.getMonteCarloWeights();
	P values = mult(floor(sub(fr, strike), 0.0), periodLength);
	values = mult(div(values, numeraire), w);
	values = div(mult(values, m.getNumeraire(pr, 0.0)), w);
	if (valuesOut) store(values, valuesOut, paths);
	if (numeraireOut) store(numeraire, numeraireOut, paths);
	if (forwardRateOut) store(fr, forwardRateOut, paths);
	return getAverage(values);
}

// Asset Bermudan option on the Black-Scholes model (BermudanOption.java, ESTIMATE_COND_EXPECTATION); returns the price
double orc_bs_bermudan_option(int seed, const double* times, int nTimes, int paths, double s0, double r, double sigma, int scheme,
		const double* exerciseDates, const double* notionals, const double* strikes, int nExercise, int numberOfBasisFunctions,
		int intrinsicValueAsBasisFunction, int useBinning, double* valuesOut, double* exerciseTimeOut, double* regressionOut) {
	BrownianMotion bm(tdFrom(times, nTimes), 1, paths, seed, 0);
	BlackScholesModel m(s0, r, sigma);
	Process pr(&m, &bm, scheme);
	BermudanOptionResult res = bermudanOptionValue(m, pr, 0.0, vecOf(exerciseDates, nExercise), vecOf(notionals, nExercise), vecOf(strikes, nExercise),
		numberOfBasisFunctions, intrinsicValueAsBasisFunction != 0, useBinning != 0);
	if (valuesOut) store(res.value, valuesOut, paths);
	if (exerciseTimeOut) store(res.exerciseTime, exerciseTimeOut, paths);
	if (regressionOut) for (size_t e = 0; e < res.regressionParameters.size(); e++)
		for (size_t k = 0; k < res.regressionParameters[e].size(); k++) regressionOut[e * numberOfBasisFunctions + k] = res.regressionParameters[e][k];
	return getAverage(res.value);
}
// localized regression: parameters for dependents y on basis b[K][n]
void orc_regression_localized(const double* basis, int K, const double* y, uint64_t n, double standardDeviations, double* xOut, double* ceOut) {
	std::vector<P> b;
	for (int k = 0; k < K; k++) b.push_back(wrap(basis + (size_t)k * n, n));
	RegressionLocalized reg(b, standardDeviations);
	P ce = reg.getConditionalExpectationLocalized(wrap(y, n));
	for (int k = 0; k < K; k++) xOut[k] = reg.lastParameters[k];
	if (ceOut) store(ce, ceOut, n);
}

// ---- CPU baselines for bench.py (bounded samples) --------------------------------------------------------
// (1) reference-shaped: the RV-op path above (one array pass + one allocation per op, single sequential MT stream).
//     Returns seconds for {Brownian generation + Euler evolution} of `paths` LMM paths.
double orc_time_lmm_reference_shaped(int seed, const double* simTimes, int nSimTimes, const double* tenorTimes, int nTenorTimes, int F,
		int paths, const double* L0, const double* sigma, const double* factorMatrix, int scheme, double* checksum) {
	auto t0 = std::chrono::steady_clock::now();
	void* hv = orc_lmm_create(seed, simTimes, nSimTimes, tenorTimes, nTenorTimes, F, paths, 0, L0, nullptr, sigma, factorMatrix, 0, 1, 1e5, scheme);
	auto* h = (LmmHandle*)hv;
	h->process->precalc();
	auto t1 = std::chrono::steady_clock::now();
	if (checksum) *checksum = getAverage(h->process->getProcessValue(nSimTimes - 1, nTenorTimes - 2));
	orc_lmm_free(hv);
	return std::chrono::duration<double>(t1 - t0).count();
}

// (2) best-effort CPU: same arithmetic (spot measure, lognormal, EULER_FUNCTIONAL or PREDICTOR_CORRECTOR family), fused per path,
//     path-parallel over `threads` host threads (each thread jumps the MT stream to its first path by skipping).
//     Writes X[t][j][p] like the reference (so the memory traffic is honest).  Returns seconds.
static void lmmFusedRange(int seed, const TimeDiscretization& sim, const TimeDiscretization& tenor, int F, int paths, int p0, int p1,
		const double* L0, const double* sigma, const double* fm, int scheme, double cap, double* X) {
	const int T = sim.getNumberOfTimeSteps(), N = tenor.getNumberOfTimeSteps();
	MersenneTwister mt((int64_t)seed);
	mt.skipWords((uint64_t)p0 * 2ull * T * F);
	std::vector<double> sq(T), dt(T), L(N), Y(N), mu(N), mu2(N), dW(F), S(F), fl((size_t)T * N * F);
	std::vector<int> first(T);
	for (int t = 0; t < T; t++) {
		dt[t] = sim.getTime(t + 1) - sim.getTime(t); sq[t] = std::sqrt(sim.getTimeStep(t));
		int f = tenor.getTimeIndex(sim.getTime(t)) + 1; if (f < 0) f = -f - 1 + 1; first[t] = f;
		for (int j = 0; j < N; j++) for (int k = 0; k < F; k++) fl[((size_t)t * N + j) * F + k] = sigma[(size_t)t * N + j] * fm[(size_t)j * F + k];
	}
	const bool functional = (scheme == EULER_FUNCTIONAL || scheme == PREDICTOR_CORRECTOR_FUNCTIONAL);
	const bool pc = (scheme == PREDICTOR_CORRECTOR || scheme == PREDICTOR_CORRECTOR_FUNCTIONAL);
	auto driftOf = [&](int t, const std::vector<double>& Lv, std::vector<double>& m) {
		for (int k = 0; k < F; k++) S[k] = 0.0;
		for (int j = first[t]; j < N; j++) {
			const double d = tenor.getTimeStep(j);
			double a = 1.0 / (Lv[j] * (d / d) + 1.0 / d);
			a = a * Lv[j];
			const double* f = &fl[((size_t)t * N + j) * F];
			for (int k = 0; k < F; k++) S[k] = S[k] + a * f[k];
			double s = S[0] * f[0] + 0.0;
			for (int k = 1; k < F; k++) s = s + S[k] * f[k];
			const double sg = sigma[(size_t)t * N + j];
			m[j] = s + (sg * sg * 1.0) * -0.5;
		}
	};
	for (int p = p0; p < p1; p++) {
		for (int j = 0; j < N; j++) { Y[j] = std::log(std::max(L0[j], 0.0)); L[j] = jmin(std::exp(Y[j]), cap); X[(size_t)j * paths + p] = L[j]; }
		for (int t = 0; t < T; t++) {
			for (int k = 0; k < F; k++) dW[k] = inverseCumulativeNormal(mt.nextDouble()) * sq[t];
			driftOf(t, L, mu);
			double* Xn = X + (size_t)(t + 1) * N * paths;
			for (int j = 0; j < first[t] && j < N; j++) Xn[(size_t)j * paths + p] = L[j];
			for (int j = first[t]; j < N; j++) {
				double y = functional ? std::log(L[j]) : Y[j];
				y = y + mu[j] * dt[t];
				const double* f = &fl[((size_t)t * N + j) * F];
				for (int k = 0; k < F; k++) y = y + dW[k] * f[k];
				Y[j] = y;
			}
			std::vector<double>& Ln = L;   // in-place is fine: drift already evaluated on the old state
			for (int j = first[t]; j < N; j++) Ln[j] = jmin(std::exp(Y[j]), cap);
			if (pc) {
				driftOf(t, L, mu2);
				for (int j = first[t]; j < N; j++) { Y[j] = Y[j] + ((mu2[j] - mu[j]) / 2.0) * dt[t]; L[j] = jmin(std::exp(Y[j]), cap); }
			}
			for (int j = first[t]; j < N; j++) Xn[(size_t)j * paths + p] = L[j];
		}
	}
}
double orc_time_lmm_fused(int seed, const double* simTimes, int nSimTimes, const double* tenorTimes, int nTenorTimes, int F, int paths,
		const double* L0, const double* sigma, const double* factorMatrix, int scheme, int threads, double* processOut /* [T+1][N][P] or NULL */) {
	TimeDiscretization sim = tdFrom(simTimes, nSimTimes), tenor = tdFrom(tenorTimes, nTenorTimes);
	const int N = nTenorTimes - 1;
	std::vector<double> local;
	double* X = processOut;
	if (!X) { local.resize((size_t)nSimTimes * N * paths); X = local.data(); }
	auto t0 = std::chrono::steady_clock::now();
	std::vector<std::thread> th;
	for (int i = 0; i < threads; i++) {
		const int p0 = (int)((int64_t)paths * i / threads), p1 = (int)((int64_t)paths * (i + 1) / threads);
		th.emplace_back(lmmFusedRange, seed, std::cref(sim), std::cref(tenor), F, paths, p0, p1, L0, sigma, factorMatrix, scheme, 1e5, X);
	}
	for (auto& t : th) t.join();
	auto t1 = std::chrono::steady_clock::now();
	return std::chrono::duration<double>(t1 - t0).count();
}

} // extern "C"
