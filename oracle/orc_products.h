// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_core.h header).  Regression estimator and the product consumers.
// J/ = /root/reference/src/main/java/net/finmath/
#pragma once
#include "orc_models.h"

namespace orc {

// Pseudo-inverse solve of a symmetric K x K system, standing in for commons-math3 3.6.1
// SingularValueDecomposition(XTX).getSolver().solve(XTy) (third-party, jar only).  For a symmetric positive
// semi-definite matrix the SVD is the eigen-decomposition; singular values <= tol = max(K*s_max*2^-52, sqrt(2^-1022))
// are dropped.  Cyclic Jacobi rotations.  "parity unpinned": no reference test pins the solve.
inline std::vector<double> solveSymmetricPseudoInverse(std::vector<double> A, const std::vector<double>& b, int K, double* condOut = nullptr) {
	std::vector<double> V((size_t)K * K, 0.0);
	for (int i = 0; i < K; i++) V[(size_t)i * K + i] = 1.0;
	for (int sweep = 0; sweep < 100; sweep++) {
		double off = 0.0;
		for (int p = 0; p < K; p++) for (int q = p + 1; q < K; q++) off += A[(size_t)p * K + q] * A[(size_t)p * K + q];
		if (off == 0.0) break;
		for (int p = 0; p < K; p++) for (int q = p + 1; q < K; q++) {
			const double apq = A[(size_t)p * K + q];
			if (apq == 0.0) continue;
			const double app = A[(size_t)p * K + p], aqq = A[(size_t)q * K + q];
			const double theta = (aqq - app) / (2.0 * apq);
			const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
			const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
			for (int k = 0; k < K; k++) {
				const double akp = A[(size_t)k * K + p], akq = A[(size_t)k * K + q];
				A[(size_t)k * K + p] = c * akp - s * akq;
				A[(size_t)k * K + q] = s * akp + c * akq;
			}
			for (int k = 0; k < K; k++) {
				const double apk = A[(size_t)p * K + k], aqk = A[(size_t)q * K + k];
				A[(size_t)p * K + k] = c * apk - s * aqk;
				A[(size_t)q * K + k] = s * apk + c * aqk;
			}
			for (int k = 0; k < K; k++) {
				const double vkp = V[(size_t)k * K + p], vkq = V[(size_t)k * K + q];
				V[(size_t)k * K + p] = c * vkp - s * vkq;
				V[(size_t)k * K + q] = s * vkp + c * vkq;
			}
		}
	}
	double smax = 0.0, smin = std::numeric_limits<double>::infinity();
	for (int i = 0; i < K; i++) { smax = std::max(smax, std::fabs(A[(size_t)i * K + i])); smin = std::min(smin, std::fabs(A[(size_t)i * K + i])); }
	if (condOut) *condOut = smax / smin;
	const double tol = std::max((double)K * smax * 0x1.0p-52, std::sqrt(0x1.0p-1022));
	std::vector<double> x(K, 0.0);
	for (int i = 0; i < K; i++) {
		const double lam = A[(size_t)i * K + i];
		if (std::fabs(lam) <= tol) continue;
		double vb = 0.0;
		for (int k = 0; k < K; k++) vb += V[(size_t)k * K + i] * b[k];
		const double w = vb / lam;
		for (int k = 0; k < K; k++) x[k] += V[(size_t)k * K + i] * w;
	}
	return x;
}

// J/montecarlo/conditionalexpectation/MonteCarloConditionalExpectationRegression.java:97-150
struct Regression {
	std::vector<P> basis;
	std::vector<double> XTX;           // cached like the solver (:125-138)
	std::vector<double> lastParameters;
	double lastCond = 0.0;
	explicit Regression(const std::vector<P>& b) { for (auto& x : b) if (x) basis.push_back(x); }
	std::vector<double> getLinearRegressionParameters(const P& y) {
		const int K = (int)basis.size();
		if (XTX.empty()) {
			XTX.assign((size_t)K * K, 0.0);
			for (int i = 0; i < K; i++) for (int j = i; j < K; j++) {
				XTX[(size_t)i * K + j] = getAverage(mult(basis[i], basis[j]));
				XTX[(size_t)j * K + i] = XTX[(size_t)i * K + j];
			}
		}
		std::vector<double> XTy(K);
		for (int i = 0; i < K; i++) XTy[i] = getAverage(mult(y, basis[i]));
		lastParameters = solveSymmetricPseudoInverse(XTX, XTy, K, &lastCond);
		return lastParameters;
	}
	P getConditionalExpectation(const P& y) {
		std::vector<double> x = getLinearRegressionParameters(y);
		P ce = mult(basis[0], x[0]);
		for (size_t i = 1; i < basis.size(); i++) ce = addProduct(ce, basis[i], x[i]);
		return ce;
	}
};

// LIBORMonteCarloSimulationFromLIBORModel, J/montecarlo/interestrate/LIBORMonteCarloSimulationFromLIBORModel.java:27-206
struct LIBORSimulation {
	LIBORMarketModel* model;
	Process* process;
	P getForwardRate(double t, double s, double e) { return model->getForwardRate(*process, t, s, e); }
	P getNumeraire(double t) { return model->getNumeraire(*process, t); }
	P getMonteCarloWeights(double) { return process->getMonteCarloWeights(); }
	P getRandomVariableForConstant(double v) { return scalar(v); }
};

// J/montecarlo/interestrate/products/Swaption.java:137-200 (discount curve = curve from the forward curve =>
// discountingAdjustment is computed from two identical curves; passed in as an array, 1.0 when equal)
inline P swaptionValue(LIBORSimulation& m, double evaluationTime, double exerciseDate, const std::vector<double>& fixingDates,
		const std::vector<double>& paymentDates, const std::vector<double>& swaprates, double notional,
		const std::vector<double>& discountingAdjustments) {
	P v = m.getRandomVariableForConstant(0.0);
	for (int period = (int)fixingDates.size() - 1; period >= 0; period--) {
		const double fixingDate = fixingDates[period], paymentDate = paymentDates[period], swaprate = swaprates[period];
		if (paymentDate <= evaluationTime) break;
		const double periodLength = paymentDate - fixingDate;
		P libor = m.getForwardRate(exerciseDate, fixingDate, paymentDate);
		P payoff = mult(mult(sub(libor, swaprate), periodLength), notional);
		const double discountingDate = std::max(fixingDate, exerciseDate);
		v = add(v, payoff);
		v = mult(discount(v, libor, paymentDate - discountingDate), discountingAdjustments.empty() ? 1.0 : discountingAdjustments[period]);
	}
	P values = floor(v, 0.0);
	values = mult(div(values, m.getNumeraire(exerciseDate)), m.getMonteCarloWeights(exerciseDate));
	values = div(mult(values, m.getNumeraire(evaluationTime)), m.getMonteCarloWeights(evaluationTime));
	return values;
}

// J/montecarlo/interestrate/products/Caplet.java:114-160 (ValueUnit.VALUE)
inline P capletValue(LIBORSimulation& m, double evaluationTime, double maturity, double periodLength, double strike,
		double daycountFraction, bool isFloorlet) {
	const double paymentDate = maturity + periodLength;
	P fr = m.getForwardRate(maturity, maturity, maturity + periodLength);
	P numeraire = m.getNumeraire(paymentDate);
	P w = m.getMonteCarloWeights(paymentDate);
	P values = !isFloorlet ? mult(floor(sub(fr, strike), 0.0), daycountFraction) : mult(cap(sub(fr, strike), 0.0), -1.0 * daycountFraction);
	values = mult(div(values, numeraire), w);
	values = div(mult(values, m.getNumeraire(evaluationTime)), m.getMonteCarloWeights(evaluationTime));
	return values;
}

// J/montecarlo/interestrate/products/BermudanSwaption.java:90-252
struct BermudanResult {
	P value, exerciseTime;
	std::vector<std::vector<double>> regressionParameters;   // per exercise date, in loop (backward) order
	std::vector<double> regressionCond;
};
inline std::vector<P> bermudanBasisFunctions(LIBORSimulation& m, double fixingDate, const std::vector<double>& fixingDates,
		const std::vector<double>& paymentDates) {
	std::vector<P> b;
	b.push_back(rvconst(NEG_INF, 1.0));                                                      // :220 new RandomVariableFromDoubleArray(1.0)
	// Arrays.binarySearch(fixingDates, fixingDate)
	int lo = 0, hi = (int)fixingDates.size() - 1, idx = -1;
	while (lo <= hi) { int mid = (lo + hi) >> 1; if (fixingDates[mid] < fixingDate) lo = mid + 1; else if (fixingDates[mid] > fixingDate) hi = mid - 1; else { idx = mid; break; } }
	if (idx < 0) idx = -(-(lo + 1));                                                         // :224-226 quirk: -index
	if (idx >= (int)fixingDates.size()) idx = (int)fixingDates.size() - 1;
	P rateShort = m.getForwardRate(fixingDate, fixingDate, paymentDates[idx]);
	P discountShort = invert(add(mult(rateShort, paymentDates[idx] - fixingDate), 1.0));
	b.push_back(discountShort);
	b.push_back(pow(discountShort, 2.0));
	P rateLong = m.getForwardRate(fixingDate, fixingDates[idx], paymentDates.back());
	P discountLong = invert(add(mult(rateLong, paymentDates.back() - fixingDates[idx]), 1.0));
	b.push_back(discountLong);
	b.push_back(pow(discountLong, 2.0));
	b.push_back(invert(m.getNumeraire(fixingDate)));
	return b;
}
// givenCoefficients (test aid): replay the induction with these regression coefficients (one vector per exercise date, in loop order)
// instead of estimating them, and weightOverride > 0 as the Monte-Carlo weight - a WINDOW of paths of a large simulation then reproduces
// the per-path values and exercise decisions of the full run (the coefficients are the only quantity that depends on all paths).
inline BermudanResult bermudanSwaptionValues(LIBORSimulation& m, double evaluationTime, const std::vector<int>& isExercise,
		const std::vector<double>& fixingDates, const std::vector<double>& periodLengths, const std::vector<double>& paymentDates,
		const std::vector<double>& notionals, const std::vector<double>& swaprates, bool isCallable,
		const std::vector<std::vector<double>>* givenCoefficients = nullptr, double weightOverride = 0.0) {
	BermudanResult res;
	size_t exerciseCount = 0;
	P values = m.getRandomVariableForConstant(0.0);
	P valuesUnderlying = m.getRandomVariableForConstant(0.0);
	P exerciseTime = m.getRandomVariableForConstant(std::numeric_limits<double>::infinity());
	for (int period = (int)fixingDates.size() - 1; period >= 0; period--) {
		const double fixingDate = fixingDates[period], exerciseDate = fixingDate, periodLength = periodLengths[period];
		const double paymentDate = paymentDates[period], notional = notionals[period], swaprate = swaprates[period];
		P libor = m.getForwardRate(fixingDate, fixingDate, fixingDate + periodLength);
		P payoff = mult(mult(sub(libor, swaprate), periodLength), notional);
		P numeraire = m.getNumeraire(paymentDate);
		P w = weightOverride > 0.0 ? scalar(weightOverride) : m.getMonteCarloWeights(paymentDate);
		payoff = mult(div(payoff, numeraire), w);
		if (isCallable) valuesUnderlying = add(valuesUnderlying, payoff); else values = add(values, payoff);
		if (isExercise[period]) {
			P trig = sub(values, valuesUnderlying);
			Regression reg(bermudanBasisFunctions(m, fixingDate, fixingDates, paymentDates));
			P triggerValues;
			if (givenCoefficients) {                                                             // :103-107 with the supplied parameters
				const std::vector<double>& x = (*givenCoefficients)[exerciseCount++];
				triggerValues = mult(reg.basis[0], x[0]);
				for (size_t i = 1; i < reg.basis.size(); i++) triggerValues = addProduct(triggerValues, reg.basis[i], x[i]);
				reg.lastParameters = x; reg.lastCond = 0.0;
			} else triggerValues = reg.getConditionalExpectation(trig);
			res.regressionParameters.push_back(reg.lastParameters);
			res.regressionCond.push_back(reg.lastCond);
			values = choose(triggerValues, values, valuesUnderlying);
			exerciseTime = choose(triggerValues, exerciseTime, scalar(exerciseDate));
		}
	}
	values = div(mult(values, m.getNumeraire(evaluationTime)), weightOverride > 0.0 ? scalar(weightOverride) : m.getMonteCarloWeights(evaluationTime));
	res.value = values;
	res.exerciseTime = exerciseTime;
	return res;
}

// J/montecarlo/assetderivativevaluation/products/EuropeanOption.java:172-193 on MonteCarloAssetModel (:85-120)
template <class Model> inline P europeanOptionValue(Model& model, Process& process, double evaluationTime, double maturity, double strike, int callOrPutSign) {
	const int ti = process.getTimeIndex(maturity);
	if (ti < 0) throw std::runtime_error("The model does not provide an interpolation of simulation time");
	P s = process.getProcessValue(ti, 0);
	P values = floor(mult(sub(s, strike), (double)callOrPutSign), 0.0);
	values = mult(div(values, model.getNumeraire(maturity)), process.getMonteCarloWeights());
	values = div(mult(values, model.getNumeraire(evaluationTime)), process.getMonteCarloWeights());
	return values;
}

// J/montecarlo/conditionalexpectation/MonteCarloConditionalExpectationRegressionLocalizedOnDependents.java:85-128
struct RegressionLocalized : Regression {
	double standardDeviations;
	RegressionLocalized(const std::vector<P>& b, double sd) : Regression(b), standardDeviations(sd) {}
	std::vector<double> getLinearRegressionParametersLocalized(P dependents) {
		P w = choose(sub(squared(dependents), std::pow(getStandardDeviation(dependents) * standardDeviations, 2.0)), scalar(0.0), scalar(1.0));
		const int K = (int)basis.size();
		std::vector<P> bl(K);
		for (int i = 0; i < K; i++) bl[i] = mult(basis[i], w);
		dependents = mult(dependents, w);
		std::vector<double> A((size_t)K * K), b(K);
		for (int i = 0; i < K; i++) for (int j = i; j < K; j++) A[(size_t)i * K + j] = A[(size_t)j * K + i] = getAverage(mult(bl[i], bl[j]));
		for (int i = 0; i < K; i++) b[i] = getAverage(mult(dependents, bl[i]));
		lastParameters = solveSymmetricPseudoInverse(A, b, K, &lastCond);
		return lastParameters;
	}
	P getConditionalExpectationLocalized(const P& y) {
		std::vector<double> x = getLinearRegressionParametersLocalized(y);
		P ce = mult(basis[0], x[0]);
		for (size_t i = 1; i < basis.size(); i++) ce = addProduct(ce, basis[i], x[i]);
		return ce;
	}
};

// J/montecarlo/assetderivativevaluation/products/BermudanOption.java:150-330, ExerciseMethod.ESTIMATE_COND_EXPECTATION
struct BermudanOptionResult { P value, exerciseTime; std::vector<std::vector<double>> regressionParameters; };
template <class Model> inline BermudanOptionResult bermudanOptionValue(Model& model, Process& process, double evaluationTime,
		const std::vector<double>& exerciseDates, const std::vector<double>& notionals, const std::vector<double>& strikes,
		int numberOfBasisFunctions, bool intrinsicValueAsBasisFunction, bool useBinning) {
	BermudanOptionResult res;
	P value = scalar(0.0);
	P exerciseTime = scalar(exerciseDates.back() + 1);
	P w = process.getMonteCarloWeights();
	for (int e = (int)exerciseDates.size() - 1; e >= 0; e--) {
		const double exerciseDate = exerciseDates[e];
		const int ti = process.getTimeIndex(exerciseDate);
		if (ti < 0) throw std::runtime_error("The model does not provide an interpolation of simulation time");
		P underlying = process.getProcessValue(ti, 0);
		P numeraire = model.getNumeraire(exerciseDate);
		P valueIfExercised = mult(div(mult(sub(underlying, strikes[e]), notionals[e]), numeraire), w);
		P bu = intrinsicValueAsBasisFunction ? floor(sub(underlying, strikes[e]), 0.0) : underlying;
		// new RandomVariableFromDoubleArray(0.0, underlying.getRealizations()) :298, :314
		std::vector<double> vals(process.getNumberOfPaths());
		for (size_t i = 0; i < vals.size(); i++) vals[i] = bu->get(i);
		P u0 = rvvec(0.0, std::vector<double>(vals));
		std::vector<P> basis;
		if (!useBinning) {
			for (int k = 0; k <= numberOfBasisFunctions - 1; k++) basis.push_back(pow(u0, (double)k));
		} else {
			std::sort(vals.begin(), vals.end());
			for (int i = 0; i < numberOfBasisFunctions; i++) {
				const double binLeft = vals[(size_t)((double)i / (double)numberOfBasisFunctions * vals.size())];
				basis.push_back(choose(sub(u0, binLeft), rvconst(NEG_INF, 1.0), rvconst(NEG_INF, 0.0)));
			}
		}
		Regression reg(basis);
		P estimated = reg.getConditionalExpectation(value);
		res.regressionParameters.push_back(reg.lastParameters);
		P exerciseCriteria = sub(estimated, valueIfExercised);
		value = choose(exerciseCriteria, value, valueIfExercised);
		exerciseTime = choose(exerciseCriteria, exerciseTime, scalar(exerciseDate));
	}
	value = div(mult(value, model.getNumeraire(evaluationTime)), w);
	res.value = value; res.exerciseTime = exerciseTime;
	return res;
}

} // namespace orc
