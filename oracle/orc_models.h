// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_core.h header).  Brownian driver, Euler scheme and the model
// callbacks, restated op-by-op on the orc::RV semantics so that rounding order follows the reference's dispatch.
// J/ = /root/reference/src/main/java/net/finmath/
#pragma once
#include "orc_core.h"
#include <map>

namespace orc {

// J/montecarlo/BrownianMotionFromMersenneRandomNumbers.java:141-191 — draw order path, time, factor; value
// [t][f][p] = ICDF(u) * sqrt(dt_t); wrapped as RandomVariableFromDoubleArray with filtration time t_{i+1}.
struct BrownianMotion {
	TimeDiscretization td;
	int F, paths, seed;
	std::vector<std::vector<P>> inc;   // [t][f]
	// pathOffset/pathCount select a shard of the single sequential stream (paths [pathOffset, pathOffset+pathCount))
	BrownianMotion(const TimeDiscretization& td_, int F_, int paths_, int seed_, long pathOffset = 0)
		: td(td_), F(F_), paths(paths_), seed(seed_) {
		const int T = td.getNumberOfTimeSteps();
		MersenneTwister mt((int64_t)seed);
		mt.skipWords((uint64_t)pathOffset * 2ull * (uint64_t)T * (uint64_t)F);
		std::vector<std::vector<std::vector<double>>> a(T, std::vector<std::vector<double>>(F, std::vector<double>(paths)));
		std::vector<double> sq(T);
		for (int t = 0; t < T; t++) sq[t] = std::sqrt(td.getTimeStep(t));
		for (int p = 0; p < paths; p++)
			for (int t = 0; t < T; t++)
				for (int f = 0; f < F; f++) {
					const double u = mt.nextDouble();
					a[t][f][p] = inverseCumulativeNormal(u) * sq[t];
				}
		inc.resize(T);
		for (int t = 0; t < T; t++)
			for (int f = 0; f < F; f++) inc[t].push_back(rvvec(td.getTime(t + 1), std::move(a[t][f])));
	}
	P getBrownianIncrement(int t, int f) const { return inc[t][f]; }
	const std::vector<P>& getIncrement(int t) const { return inc[t]; }
	P getRandomVariableForConstant(double v) const { return scalar(v); }   // default factory -> Scalar
};

struct Process;

// J/montecarlo/model/ProcessModel.java:47-174
struct ProcessModel {
	virtual ~ProcessModel() {}
	virtual int getNumberOfComponents() const = 0;
	virtual std::vector<P> getInitialState(const Process&) = 0;
	virtual std::vector<P> getDrift(const Process&, int timeIndex, const std::vector<P>& x) = 0;   // null entries = frozen
	virtual std::vector<P> getFactorLoading(const Process&, int timeIndex, int component, const std::vector<P>& x) = 0;
	virtual P applyStateSpaceTransform(int timeIndex, int component, const P& y) = 0;
	virtual bool hasInverse() const = 0;
	virtual P applyStateSpaceTransformInverse(int timeIndex, int component, const P& x) = 0;
};

enum Scheme { EULER = 0, PREDICTOR_CORRECTOR = 1, EULER_FUNCTIONAL = 2, PREDICTOR_CORRECTOR_FUNCTIONAL = 3 };

// J/montecarlo/process/EulerSchemeFromProcessModel.java:170-326
struct Process {
	ProcessModel* model;
	const BrownianMotion* bm;
	int scheme;
	std::vector<std::vector<P>> X;     // [timeIndex][component]
	P weights;
	Process(ProcessModel* m, const BrownianMotion* b, int scheme_ = -1) : model(m), bm(b) {
		scheme = scheme_ >= 0 ? scheme_ : (m->hasInverse() ? EULER_FUNCTIONAL : EULER);   // :108-123
	}
	const TimeDiscretization& getTimeDiscretization() const { return bm->td; }
	double getTime(int i) const { return bm->td.getTime(i); }
	int getTimeIndex(double t) const { return bm->td.getTimeIndex(t); }
	int getNumberOfPaths() const { return bm->paths; }
	P getMonteCarloWeights() { precalc(); return weights; }
	P getProcessValue(int timeIndex, int component) { precalc(); return X.at(timeIndex).at(component); }

	void precalc() {
		if (!X.empty()) return;
		const int T = bm->td.getNumberOfTimeSteps();
		const int N = model->getNumberOfComponents();
		X.assign(T + 1, std::vector<P>(N));
		weights = bm->getRandomVariableForConstant(1.0 / bm->paths);                       // :184
		std::vector<P> cur = model->getInitialState(*this);
		for (int c = 0; c < N; c++) X[0][c] = model->applyStateSpaceTransform(0, c, cur[c]);
		for (int ti = 1; ti <= T; ti++) {
			const double deltaT = getTime(ti) - getTime(ti - 1);
			std::vector<P> drift = model->getDrift(*this, ti - 1, X[ti - 1]);
			const std::vector<P>& dW = bm->getIncrement(ti - 1);
			for (int c = 0; c < N; c++) {
				if (!drift[c]) { X[ti][c] = X[ti - 1][c]; continue; }                         // :236-240, :285
				if (scheme == EULER_FUNCTIONAL || scheme == PREDICTOR_CORRECTOR_FUNCTIONAL)
					cur[c] = model->applyStateSpaceTransformInverse(ti - 1, c, X[ti - 1][c]);
				std::vector<P> fl = model->getFactorLoading(*this, ti - 1, c, X[ti - 1]);
				if (fl.empty()) { X[ti][c] = X[ti - 1][c]; continue; }
				cur[c] = addProduct(cur[c], drift[c], deltaT);
				cur[c] = addSumProduct(cur[c], fl, dW);
				X[ti][c] = model->applyStateSpaceTransform(ti, c, cur[c]);
			}
			if (scheme == PREDICTOR_CORRECTOR || scheme == PREDICTOR_CORRECTOR_FUNCTIONAL) {   // :292-314
				std::vector<P> driftP = model->getDrift(*this, ti - 1, X[ti]);
				for (int c = 0; c < N; c++) {
					if (!driftP[c] || !drift[c]) continue;
					P adj = mult(div(sub(driftP[c], drift[c]), 2.0), deltaT);
					cur[c] = add(cur[c], adj);
					X[ti][c] = model->applyStateSpaceTransform(ti, c, cur[c]);
				}
			}
		}
	}
};

// J/montecarlo/assetderivativevaluation/models/BlackScholesModel.java:60-139
struct BlackScholesModel : ProcessModel {
	P initialValue, riskFreeRate, volatility;
	std::vector<P> initialState, drift, fl;
	BlackScholesModel(double s0, double r, double sigma) {
		initialValue = scalar(s0); riskFreeRate = scalar(r); volatility = scalar(sigma);
		initialState = { log(initialValue) };
		drift = { sub(riskFreeRate, div(squared(volatility), 2)) };
		fl = { volatility };
	}
	int getNumberOfComponents() const override { return 1; }
	std::vector<P> getInitialState(const Process&) override { return initialState; }
	std::vector<P> getDrift(const Process&, int, const std::vector<P>&) override { return drift; }
	std::vector<P> getFactorLoading(const Process&, int, int, const std::vector<P>&) override { return fl; }
	P applyStateSpaceTransform(int, int, const P& y) override { return exp(y); }
	bool hasInverse() const override { return true; }
	P applyStateSpaceTransformInverse(int, int, const P& x) override { return log(x); }
	P getNumeraire(double time) const { return exp(mult(riskFreeRate, time)); }
};

// J/montecarlo/assetderivativevaluation/models/HestonModel.java:325-430 (constant-rate constructor path)
struct HestonModel : ProcessModel {
	enum HScheme { REFLECTION = 0, FULL_TRUNCATION = 1 };
	P initialValue, riskFreeRate, volatility, discountRate, theta, kappa, xi, rho, rhoBar, ZERO;
	int hscheme;
	HestonModel(double s0, double r, double sigma, double discRate, double theta_, double kappa_, double xi_, double rho_, int hs) {
		initialValue = scalar(s0); riskFreeRate = scalar(r); volatility = scalar(sigma); discountRate = scalar(discRate);
		theta = scalar(theta_); kappa = scalar(kappa_); xi = scalar(xi_); rho = scalar(rho_);
		rhoBar = sqrt(mult(sub(squared(rho), 1), -1));                                    // :182
		ZERO = scalar(0.0);
		hscheme = hs;
	}
	int getNumberOfComponents() const override { return 2; }
	std::vector<P> getInitialState(const Process&) override { return { log(initialValue), squared(volatility) }; }
	P truncated(const P& v) const { return hscheme == FULL_TRUNCATION ? floor(v, 0.0) : abs(v); }
	std::vector<P> getDrift(const Process&, int, const std::vector<P>& x) override {
		P var = truncated(x[1]);
		return { sub(riskFreeRate, div(var, 2.0)), mult(sub(theta, var), kappa) };          // :361-362
	}
	std::vector<P> getFactorLoading(const Process&, int, int c, const std::vector<P>& x) override {
		P vol = sqrt(truncated(x[1]));
		if (c == 0) return { vol, ZERO };
		P v = mult(vol, xi);
		return { mult(v, rho), mult(v, rhoBar) };
	}
	P applyStateSpaceTransform(int, int c, const P& y) override { return c == 0 ? exp(y) : y; }
	bool hasInverse() const override { return true; }
	P applyStateSpaceTransformInverse(int, int c, const P& x) override { return c == 0 ? log(x) : x; }
	P getNumeraire(double time) const { return exp(mult(discountRate, time)); }
};

// J/montecarlo/interestrate/models/LIBORMarketModelFromCovarianceModel.java.  The covariance model is the
// deterministic table pair (sigma[t][j], F[j][k]) of LIBORCovarianceModelFromVolatilityAndCorrelation (:47-93):
// factor loading = sigma.mult(F[j][k]); variance = sigma.mult(sigma).mult(corr_jj) with corr_jj == 1.0.
struct LIBORMarketModel : ProcessModel {
	enum Measure { SPOT = 0, TERMINAL = 1 };
	enum StateSpace { NORMAL = 0, LOGNORMAL = 1 };
	TimeDiscretization tenor;          // liborPeriodDiscretization
	std::vector<double> L0;            // forward curve values L_j(0)
	std::vector<double> discountFactors; // P^d(T_i) on the tenor grid (size N+1) or empty (no discount curve)
	std::vector<double> sigma;         // [T][N] instantaneous volatilities on the simulation grid
	std::vector<double> factorMatrix;  // [N][F]
	int F;
	int measure = SPOT, stateSpace = LOGNORMAL;
	double liborCap = 1e5;
	// caches (:202-204)
	std::map<int, P> numeraires;
	std::map<int, P> numeraireDiscountFactors;

	int getNumberOfComponents() const override { return tenor.getNumberOfTimeSteps(); }
	int getLiborPeriodIndex(double time) const { return tenor.getTimeIndex(time); }
	double getLiborPeriod(int i) const { return tenor.getTime(i); }

	std::vector<P> getInitialState(const Process&) override {                               // :1080-1093
		std::vector<P> s;
		for (int j = 0; j < getNumberOfComponents(); j++) {
			const double rate = L0[j];
			s.push_back(scalar(stateSpace == LOGNORMAL ? std::log(std::max(rate, 0.0)) : rate));
		}
		return s;
	}
	int simTimeIndexOf(const Process& p, double time) const {                              // AbstractLIBORCovarianceModel.java:58-63
		int ti = p.getTimeIndex(time);
		if (ti < 0) ti = -ti - 2;
		return ti;
	}
	P volatility(int ti, int j) const { return scalar(sigma[(size_t)ti * getNumberOfComponents() + j]); }
	std::vector<P> factorLoadingAt(int ti, int j) const {                                   // ...VolatilityAndCorrelation.java:47-57
		std::vector<P> fl(F);
		P vol = volatility(ti, j);
		for (int k = 0; k < F; k++) fl[k] = mult(vol, factorMatrix[(size_t)j * F + k]);
		return fl;
	}
	std::vector<P> getFactorLoading(const Process& p, int timeIndex, int c, const std::vector<P>&) override {   // :1193-1197
		return factorLoadingAt(simTimeIndexOf(p, p.getTime(timeIndex)), c);
	}
	std::vector<P> getDrift(const Process& p, int timeIndex, const std::vector<P>& x) override {   // :1124-1191
		const double time = p.getTime(timeIndex);
		int first = getLiborPeriodIndex(time) + 1;
		if (first < 0) first = -first - 1 + 1;
		const int N = getNumberOfComponents();
		P zero = scalar(0.0);
		std::vector<P> drift(N);
		for (int c = first; c < N; c++) drift[c] = zero;
		std::vector<P> sums(F, zero);
		const int ti = simTimeIndexOf(p, time);
		if (measure == SPOT) {
			for (int c = first; c < N; c++) {
				const double pl = tenor.getTimeStep(c);
				P fr = x[c];
				P ost = discount(scalar(pl), fr, pl);
				if (stateSpace == LOGNORMAL) ost = mult(ost, fr);
				std::vector<P> fl = factorLoadingAt(ti, c);
				for (int k = 0; k < F; k++) sums[k] = addProduct(sums[k], ost, fl[k]);
				drift[c] = addSumProduct(drift[c], sums, fl);
			}
		} else {
			for (int c = N - 1; c >= first; c--) {
				const double pl = tenor.getTimeStep(c);
				P fr = x[c];
				P ost = discount(scalar(-pl), fr, pl);
				if (stateSpace == LOGNORMAL) ost = mult(ost, fr);
				std::vector<P> fl = factorLoadingAt(ti, c);
				drift[c] = addSumProduct(drift[c], sums, fl);
				for (int k = 0; k < F; k++) sums[k] = addProduct(sums[k], ost, fl[k]);
			}
		}
		if (stateSpace == LOGNORMAL) {
			for (int c = first; c < N; c++) {
				P vol = volatility(ti, c);
				P variance = mult(mult(vol, vol), 1.0);
				drift[c] = addProduct(drift[c], variance, -0.5);
			}
		}
		return drift;
	}
	P applyStateSpaceTransform(int, int, const P& y) override {                             // :1199-1212
		P v = y;
		if (stateSpace == LOGNORMAL) v = exp(v);
		if (!std::isinf(liborCap)) v = cap(v, liborCap);
		return v;
	}
	bool hasInverse() const override { return true; }
	P applyStateSpaceTransformInverse(int, int, const P& x) override { return stateSpace == LOGNORMAL ? log(x) : x; }

	P getLIBOR(Process& p, int timeIndex, int liborIndex) { return p.getProcessValue(timeIndex, liborIndex); }

	// Interpolation of forward rates on fractional tenor points (:1244-1281, :1324-1395).  interpolationMethod: 0 LINEAR,
	// 1 LOG_LINEAR_UNCORRECTED (the default, :186-194); LOG_LINEAR_CORRECTED is not restated.
	int interpolationMethod = 1;
	// getForwardRateCurve().getForward(model, time, paymentOffset): the curve is a host-side input given on the tenor grid (the curve
	// classes are outside the path); between grid points the forward of the tenor period containing `time` is used -
	// for a flat curve (every configuration here) that is what the reference's ForwardCurveInterpolation returns as well.
	double analyticForward(double time) const {
		int i = tenor.getTimeIndex(time);
		if (i < 0) i = -i - 2;
		i = std::max(0, std::min(i, getNumberOfComponents() - 1));
		return L0[i];
	}
	P getOnePlusInterpolatedLIBORDt(Process& p, int timeIndex, double periodStartTime, int liborPeriodIndex) {     // :1324-1395
		const double tenorPeriodStartTime = getLiborPeriod(liborPeriodIndex), tenorPeriodEndTime = getLiborPeriod(liborPeriodIndex + 1);
		const double tenorDt = tenorPeriodEndTime - tenorPeriodStartTime;
		if (tenorPeriodStartTime < p.getTime(timeIndex)) {
			timeIndex = std::min(timeIndex, p.getTimeIndex(tenorPeriodStartTime));
			if (timeIndex < 0) throw std::invalid_argument("Tenor discretization not part of time discretization.");
		}
		P onePlusLongLIBORDt = add(mult(getLIBOR(p, timeIndex, liborPeriodIndex), tenorDt), 1.0);
		const double smallDt = tenorPeriodEndTime - periodStartTime;
		const double alpha = smallDt / tenorDt;
		P onePlusInterpolatedLIBORDt;
		if (interpolationMethod == 0) onePlusInterpolatedLIBORDt = add(mult(onePlusLongLIBORDt, alpha), 1 - alpha);
		else if (interpolationMethod == 1) onePlusInterpolatedLIBORDt = exp(mult(log(onePlusLongLIBORDt), alpha));
		else throw std::runtime_error("oracle: LOG_LINEAR_CORRECTED interpolation not restated");
		const double analyticOnePlusLongLIBORDt = 1 + analyticForward(tenorPeriodStartTime) * tenorDt;
		const double analyticOnePlusShortLIBORDt = 1 + analyticForward(periodStartTime) * smallDt;
		const double analyticOnePlusInterpolatedLIBORDt = interpolationMethod == 0 ? analyticOnePlusLongLIBORDt * alpha + (1 - alpha)
		                                                                           : std::exp(std::log(analyticOnePlusLongLIBORDt) * alpha);
		return mult(onePlusInterpolatedLIBORDt, analyticOnePlusShortLIBORDt / analyticOnePlusInterpolatedLIBORDt);
	}
	// :1237-1305
	P getForwardRate(Process& p, double time, double periodStart, double periodEnd) {
		const int ps = getLiborPeriodIndex(periodStart), pe = getLiborPeriodIndex(periodEnd);
		time = std::min(time, periodStart);
		int ti = p.getTimeIndex(time);
		if (ti < 0) {
			ti = -ti - 2;
			if (time - p.getTime(ti) > p.getTime(ti + 1) - time) ti++;                        // ROUND_NEAREST
		}
		if (pe < 0) {                                                                         // :1257-1264
			const int previousEndIndex = (-pe - 1) - 1;
			const double nextEndTime = getLiborPeriod(previousEndIndex + 1);
			P onePlusLongLIBORdt = add(mult(getForwardRate(p, time, periodStart, nextEndTime), nextEndTime - periodStart), 1.0);
			P onePlusInterpolatedLIBORDt = getOnePlusInterpolatedLIBORDt(p, ti, periodEnd, previousEndIndex);
			return div(orc::sub(div(onePlusLongLIBORdt, onePlusInterpolatedLIBORDt), 1.0), periodEnd - periodStart);
		}
		if (ps < 0) {                                                                         // :1267-1279
			const int previousStartIndex = (-ps - 1) - 1;
			const double nextStartTime = getLiborPeriod(previousStartIndex + 1);
			if (nextStartTime > periodEnd) throw std::runtime_error("Interpolation not possible.");
			if (nextStartTime == periodEnd) return div(orc::sub(getOnePlusInterpolatedLIBORDt(p, ti, periodStart, previousStartIndex), 1.0), periodEnd - periodStart);
			P onePlusLongLIBORdt = add(mult(getForwardRate(p, time, nextStartTime, periodEnd), periodEnd - nextStartTime), 1.0);
			P onePlusInterpolatedLIBORDt = getOnePlusInterpolatedLIBORDt(p, ti, periodStart, previousStartIndex);
			return div(orc::sub(mult(onePlusLongLIBORdt, onePlusInterpolatedLIBORDt), 1.0), periodEnd - periodStart);
		}
		if (ps + 1 == pe) return getLIBOR(p, ti, ps);
		P acc;
		for (int k = ps; k < pe; k++) {
			const double sub = getLiborPeriod(k + 1) - getLiborPeriod(k);
			P l = getLIBOR(p, ti, k);
			acc = !acc ? add(mult(l, sub), 1.0) : accrue(acc, l, sub);
		}
		return div(orc::sub(acc, 1.0), periodEnd - periodStart);
	}
	// :1017-1074 (spot / terminal on the tenor grid)
	P numeraireUnadjustedAtIndex(Process& p, int li) {
		auto it = numeraires.find(li);
		if (it != numeraires.end()) return it->second;
		P n;
		if (measure == TERMINAL) {
			int ti = p.getTimeIndex(tenor.getTime(li));
			if (ti < 0) ti = -ti - 1;
			n = scalar(1.0);
			for (int k = li; k <= tenor.getNumberOfTimeSteps() - 1; k++) n = discount(n, getLIBOR(p, ti, k), tenor.getTimeStep(k));
		} else {
			if (li != 0) {
				int ti = p.getTimeIndex(tenor.getTime(li - 1));
				if (ti < 0) ti = -ti - 1;
				n = accrue(numeraireUnadjustedAtIndex(p, li - 1), getLIBOR(p, ti, li - 1), tenor.getTimeStep(li - 1));
			} else n = scalar(1.0);
		}
		numeraires[li] = n;
		return n;
	}
	P numeraireUnadjusted(Process& p, double time) {                                        // :962-1015
		const int li = getLiborPeriodIndex(time);
		if (li < 0) {                                                                       // :969-1006
			const int upperIndex = -li - 1, lowerIndex = upperIndex - 1;
			if (lowerIndex < 0) throw std::invalid_argument("Numeraire requested for a time before the tenor grid. Unsupported");
			P n;
			if (measure == TERMINAL) {
				n = scalar(1.0);
				for (int k = upperIndex; k <= tenor.getNumberOfTimeSteps() - 1; k++)
					n = discount(n, getLIBOR(p, p.getTimeIndex(std::min(time, tenor.getTime(k))), k), tenor.getTimeStep(k));
			} else n = numeraireUnadjusted(p, getLiborPeriod(upperIndex));
			return discount(n, getForwardRate(p, time, time, getLiborPeriod(upperIndex)), getLiborPeriod(upperIndex) - time);
		}
		return numeraireUnadjustedAtIndex(p, li);
	}
	P defaultableZeroBondAsOfTimeZeroAt(double time) {                                      // :886-905 (interpolation on the tenor grid)
		const int timeIndex = tenor.getTimeIndex(time);
		if (timeIndex >= 0) return defaultableZeroBondAsOfTimeZero(timeIndex);
		const int timeIndexPrev = std::min(-timeIndex - 2, tenor.getNumberOfTimes() - 2), timeIndexNext = timeIndexPrev + 1;
		const double timePrev = tenor.getTime(timeIndexPrev), timeNext = tenor.getTime(timeIndexNext);
		P prev = defaultableZeroBondAsOfTimeZero(timeIndexPrev), next = defaultableZeroBondAsOfTimeZero(timeIndexNext);
		return mult(prev, pow(div(next, prev), (time - timePrev) / (timeNext - timePrev)));
	}
	P defaultableZeroBondAsOfTimeZero(int timeIndex) {                                      // :915-944
		if (numeraireDiscountFactors.empty()) {
			P adj = scalar(discountFactors[0]);
			numeraireDiscountFactors[0] = adj;
			for (int i = 0; i < tenor.getNumberOfTimeSteps(); i++) {
				const double dfPrev = discountFactors[i], dfNext = discountFactors[i + 1], ts = tenor.getTimeStep(i);
				P fr = scalar((dfPrev / dfNext - 1.0) / ts);
				adj = discount(adj, fr, ts);
				numeraireDiscountFactors[i + 1] = adj;
			}
		}
		return numeraireDiscountFactors.at(timeIndex);
	}
	P getNumeraire(Process& p, double time) {                                               // :859-876
		if (time < 0) throw std::runtime_error("oracle: numeraire for negative time not restated");
		P n = numeraireUnadjusted(p, time);
		if (!discountFactors.empty()) {
			P dz = defaultableZeroBondAsOfTimeZeroAt(time);
			const double nonDefaultableZeroBond = getAverage(mult(invert(n), numeraireUnadjusted(p, 0.0)));
			n = div(mult(n, nonDefaultableZeroBond), dz);
		}
		return n;
	}
};

// J/montecarlo/interestrate/models/HullWhiteModel.java:277-424 (process callbacks) with the piecewise-constant closed forms
// :584-795 (getMRTime, getB, getV, getDV).  Volatility model = ShortRateVolatilityModelAsGiven (Scalars).
struct HullWhiteModel : ProcessModel {
	TimeDiscretization volTimes;       // time discretization of the volatility model
	std::vector<double> vol, mr;       // per volatility-time index

	int volIndex(double t) const { int i = volTimes.getTimeIndex(t); if (i < 0) i = -i - 2; return i; }
	P meanReversion(int i) const { return scalar(mr.at(i)); }
	P volatility(int i) const { return scalar(vol.at(i)); }

	P getMRTime(double time, double maturity) const {                                       // :584-607
		const int i0 = volIndex(time), i1 = volIndex(maturity);
		P integral = scalar(0.0);
		double timePrev = time, timeNext;
		for (int ti = i0 + 1; ti <= i1; ti++) {
			timeNext = volTimes.getTime(ti);
			integral = add(integral, mult(meanReversion(ti - 1), timeNext - timePrev));
			timePrev = timeNext;
		}
		timeNext = maturity;
		integral = add(integral, mult(meanReversion(i1), timeNext - timePrev));
		return integral;
	}
	P getB(double time, double maturity) const {                                            // :609-640
		const int i0 = volIndex(time), i1 = volIndex(maturity);
		P integral = scalar(0.0);
		double timePrev = time, timeNext;
		for (int ti = i0 + 1; ti <= i1; ti++) {
			timeNext = volTimes.getTime(ti);
			integral = add(integral, div(sub(exp(mult(getMRTime(timeNext, maturity), -1.0)), exp(mult(getMRTime(timePrev, maturity), -1.0))), meanReversion(ti - 1)));
			timePrev = timeNext;
		}
		timeNext = maturity;
		integral = add(integral, div(sub(exp(mult(getMRTime(timeNext, maturity), -1.0)), exp(mult(getMRTime(timePrev, maturity), -1.0))), meanReversion(i1)));
		return integral;
	}
	P vTerm(const P& v2mr2, const P& eNext, const P& ePrev, const P& mrv, double dt) const {
		return mult(v2mr2, add(add(div(mult(sub(eNext, ePrev), -2), mrv), div(div(sub(squared(eNext), squared(ePrev)), mrv), 2.0)), dt));
	}
	P getV(double time, double maturity) const {                                            // :642-690
		if (time == maturity) return scalar(0.0);
		const int i0 = volIndex(time), i1 = volIndex(maturity);
		P integral = scalar(0.0);
		double timePrev = time, timeNext;
		P ePrev = exp(mult(getMRTime(timePrev, maturity), -1));
		for (int ti = i0 + 1; ti <= i1; ti++) {
			timeNext = volTimes.getTime(ti);
			P m = meanReversion(ti - 1), v = volatility(ti - 1);
			P eNext = exp(mult(getMRTime(timeNext, maturity), -1));
			integral = add(integral, vTerm(div(squared(v), squared(m)), eNext, ePrev, m, timeNext - timePrev));
			timePrev = timeNext; ePrev = eNext;
		}
		timeNext = maturity;
		P m = meanReversion(i1), v = volatility(i1);
		P eNext = exp(mult(getMRTime(timeNext, maturity), -1));
		return add(integral, vTerm(div(squared(v), squared(m)), eNext, ePrev, m, timeNext - timePrev));
	}
	P getDV(double time, double maturity) const {                                           // :692-738
		if (time == maturity) return scalar(0.0);
		const int i0 = volIndex(time), i1 = volIndex(maturity);
		P integral = scalar(0.0);
		double timePrev = time, timeNext;
		P ePrev = exp(mult(getMRTime(timePrev, maturity), -1));
		auto term = [](const P& v2mr2, const P& eNext, const P& ePrev) {
			return mult(v2mr2, add(sub(eNext, ePrev), div(sub(squared(eNext), squared(ePrev)), -2.0)));
		};
		for (int ti = i0 + 1; ti <= i1; ti++) {
			timeNext = volTimes.getTime(ti);
			P m = meanReversion(ti - 1), v = volatility(ti - 1);
			P eNext = exp(mult(getMRTime(timeNext, maturity), -1));
			integral = add(integral, term(div(squared(v), squared(m)), eNext, ePrev));
			timePrev = timeNext; ePrev = eNext;
		}
		timeNext = maturity;
		P m = meanReversion(i1), v = volatility(i1);
		P eNext = exp(mult(getMRTime(timeNext, maturity), -1));
		return add(integral, term(div(squared(v), squared(m)), eNext, ePrev));
	}

	int getNumberOfComponents() const override { return 2; }
	std::vector<P> getInitialState(const Process&) override { P z = scalar(0.0); return { z, z }; }   // :298-303
	std::vector<P> getDrift(const Process& p, int timeIndex, const std::vector<P>& x) override {       // :367-387
		const double time = p.getTime(timeIndex), timeNext = p.getTime(timeIndex + 1);
		if (timeNext == time) return { P(), P() };
		P m = meanReversion(volIndex(time));
		P d0 = mult(x[0], mult(m, div(getB(time, timeNext), -1 * (timeNext - time))));
		P d1 = mult(x[0], div(getB(time, timeNext), timeNext - time));
		return { d0, d1 };
	}
	std::vector<P> getFactorLoading(const Process& p, int timeIndex, int c, const std::vector<P>&) override {   // :389-424
		const double time = p.getTime(timeIndex), timeNext = p.getTime(timeIndex + 1);
		const int vi = volIndex(time);
		P m = meanReversion(vi);
		P mrt = mult(m, -2.0 * (timeNext - time));
		P scaling = sqrt(div(sub(exp(mrt), 1.0), mrt));
		P volEff = mult(scaling, volatility(vi));
		if (c == 0) return { volEff, scalar(0.0) };
		P volLogNum = sqrt(div(getV(time, timeNext), timeNext - time));
		P rho = div(div(getDV(time, timeNext), timeNext - time), mult(volEff, volLogNum));
		return { mult(volLogNum, rho), mult(volLogNum, sqrt(mult(sub(squared(rho), 1), -1))) };
	}
	P applyStateSpaceTransform(int, int, const P& y) override { return y; }
	bool hasInverse() const override { return true; }
	P applyStateSpaceTransformInverse(int, int, const P& x) override { return x; }

	// ---- term structure functions (:305-357, :431-582, :797-950).  Curves are host inputs given on the curve time grid:
	// dfDiscount[i] = discountCurve.getDiscountFactor(T_i), dfForward[i] = discountCurveFromForwardCurve.getDiscountFactor(T_i).
	TimeDiscretization curveTimes;     // liborPeriodDiscretization (isInterpolateDiscountFactorsOnLiborPeriodDiscretization = true)
	std::vector<double> dfDiscount, dfForward;
	std::vector<P> numeraireDiscountFactors, dfFromForwardCache, forwardRateCache;

	P getShortRateConditionalVariance(double time, double maturity) const {                    // :740-775
		const int i0 = volIndex(time), i1 = volIndex(maturity);
		P integral = scalar(0.0);
		double timePrev = time, timeNext;
		P ePrev = exp(mult(getMRTime(timePrev, maturity), -2));
		for (int ti = i0 + 1; ti <= i1; ti++) {
			timeNext = volTimes.getTime(ti);
			P m = meanReversion(ti - 1), v = volatility(ti - 1);
			P eNext = exp(mult(getMRTime(timeNext, maturity), -2));
			integral = add(integral, mult(div(squared(v), m), div(sub(eNext, ePrev), 2)));
			timePrev = timeNext; ePrev = eNext;
		}
		timeNext = maturity;
		P m = meanReversion(i1), v = volatility(i1);
		P eNext = exp(mult(getMRTime(timeNext, maturity), -2));
		return add(integral, mult(div(squared(v), m), div(sub(eNext, ePrev), 2)));
	}
	P forwardRateInitialValue(int i) {                                                         // :935-950
		while ((int)forwardRateCache.size() <= i) {
			const int k = (int)forwardRateCache.size();
			forwardRateCache.push_back(scalar((dfForward[k] / dfForward[k + 1] - 1.0) / curveTimes.getTimeStep(k)));
		}
		return forwardRateCache[i];
	}
	P dfFromForwardCurveAt(int timeIndex) {                                                    // :912-933
		while ((int)dfFromForwardCache.size() <= timeIndex) {
			const int i = (int)dfFromForwardCache.size();
			if (i == 0) dfFromForwardCache.push_back(scalar(dfForward[0]));
			else dfFromForwardCache.push_back(div(dfFromForwardCache[i - 1], add(mult(forwardRateInitialValue(i - 1), curveTimes.getTimeStep(i - 1)), 1.0)));
		}
		return dfFromForwardCache[timeIndex];
	}
	P curveInterp(double time, const std::function<P(int)>& at) {                              // :845-860, :896-910
		const int ti = curveTimes.getTimeIndex(time);
		if (ti >= 0) return at(ti);
		const int prev = std::min(-ti - 2, curveTimes.getNumberOfTimes() - 2), next = prev + 1;
		const double tp = curveTimes.getTime(prev), tn = curveTimes.getTime(next);
		P a = at(prev), b = at(next);
		return mult(a, pow(div(b, a), (time - tp) / (tn - tp)));
	}
	P dfFromForwardCurve(double time) { return curveInterp(time, [this](int i) { return dfFromForwardCurveAt(i); }); }
	P discountFactorAt(int timeIndex) {                                                        // :862-889
		if (numeraireDiscountFactors.empty()) {
			P adj = scalar(dfDiscount[0]);
			numeraireDiscountFactors.push_back(adj);
			for (int i = 0; i < curveTimes.getNumberOfTimeSteps(); i++) {
				const double ts = curveTimes.getTimeStep(i);
				adj = discount(adj, scalar((dfDiscount[i] / dfDiscount[i + 1] - 1.0) / ts), ts);
				numeraireDiscountFactors.push_back(adj);
			}
		}
		return numeraireDiscountFactors.at(timeIndex);
	}
	P discountFactor(double time) { return curveInterp(time, [this](int i) { return discountFactorAt(i); }); }
	P zeroRateFromForwardCurve(double time) {                                                  // :891-902 (same index twice: sic)
		int ti = curveTimes.getTimeIndex(time);
		if (ti < 0) ti = std::min(-ti - 2, curveTimes.getNumberOfTimes() - 2);
		return div(log(div(dfFromForwardCurveAt(ti), dfFromForwardCurveAt(ti))), curveTimes.getTimeStep(ti));
	}
	P getShortRate(Process& p, int timeIndex) {                                                // :493-510
		const double time = p.getTime(timeIndex);
		P value = add(p.getProcessValue(timeIndex, 0), getDV(0, time));
		return add(value, zeroRateFromForwardCurve(time));
	}
	P getA(Process& p, double time, double maturity) {                                         // :543-557
		P zeroRate = zeroRateFromForwardCurve(time);
		P forwardBond = log(div(dfFromForwardCurve(maturity), dfFromForwardCurve(time)));
		P B = getB(time, maturity);
		P lnA = add(sub(mult(B, zeroRate), mult(squared(B), div(getShortRateConditionalVariance(0, time), 2))), forwardBond);
		return exp(lnA);
	}
	P getZeroCouponBond(Process& p, double time, double maturity) {                            // :512-523
		const int ti = p.getTimeIndex(time);
		if (ti < 0) {
			const double timeLo = p.getTime(-ti - 1 - 1);
			return div(getZeroCouponBond(p, timeLo, maturity), getZeroCouponBond(p, timeLo, time));
		}
		return mult(exp(mult(getShortRate(p, ti), mult(getB(time, maturity), -1))), getA(p, time, maturity));
	}
	P getForwardRate(Process& p, double time, double periodStart, double periodEnd) {          // :431-435
		return div(sub(div(getZeroCouponBond(p, time, periodStart), getZeroCouponBond(p, time, periodEnd)), 1.0), periodEnd - periodStart);
	}
	P getNumeraire(Process& p, double time) {                                                  // :305-357
		if (time == p.getTime(0)) return scalar(1.0);
		const int ti = p.getTimeIndex(time);
		if (ti < 0) {                                                                          // :317-333 log-linear interpolation
			const int previousTimeIndex = (-ti - 1) - 1;
			const double previousTime = p.getTime(previousTimeIndex), nextTime = p.getTime(previousTimeIndex + 1);
			return exp(div(add(mult(log(getNumeraire(p, previousTime)), nextTime - time), mult(log(getNumeraire(p, nextTime)), time - previousTime)),
			               nextTime - previousTime));
		}
		P logNum = add(p.getProcessValue(ti, 1), mult(getV(0, time), 0.5));
		P n = exp(logNum);
		n = mult(n, getAverage(invert(n)));
		P df = dfDiscount.empty() ? dfFromForwardCurve(time)
			: mult(div(discountFactor(time), getAverage(dfFromForwardCurve(time))), dfFromForwardCurve(time));
		return div(n, df);
	}
};

} // namespace orc
