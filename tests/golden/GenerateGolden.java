/*
 * GenerateGolden.java — emits REFERENCE-HELD fixtures for the parity tests: numbers computed by finmath-lib itself (its own CPU path,
 * RandomVariableFromDoubleArray), written as tests/golden/ref_*.json.  This image has no JDK, so the oracle (oracle/) is pinned only
 * against independent sources ("parity unpinned" against the reference itself, DESIGN.md §2); anyone with a JDK and finmath-lib on the
 * class path turns that into reference-held pinning with
 *
 *     javac -cp finmath-lib-6.0.x.jar:commons-math3-3.6.1.jar:. tests/golden/GenerateGolden.java -d /tmp/golden
 *     java  -cp finmath-lib-6.0.x.jar:commons-math3-3.6.1.jar:commons-lang3-3.x.jar:/tmp/golden GenerateGolden tests/golden
 *
 * and commits the ref_*.json files; tests/test_cpu_oracle.py::test_reference_held_fixtures picks them up (and is skipped, saying why,
 * while they are absent).  Doubles are written as their IEEE-754 bit patterns (hex) so that nothing is lost in decimal printing.
 *
 * What is emitted (the §8c list: uniforms, increments, a C1 / C4 path window, the C4 swaption and C5 Bermudan prices):
 *   ref_mt.json        MersenneTwister(seed).nextDouble() x 64 for seeds 3141, 31415, 53252, -1; AS241 of the first 64 uniforms of 3141
 *   ref_brownian.json  BrownianMotionFromMersenneRandomNumbers(dt 0.5, 4 steps, 3 factors, 16 paths, seed 3141): all increments
 *   ref_bs.json        C1 shape at 1000 paths: S(T) of 16 paths, the call price
 *   ref_lmm.json       C4 set-up (40 rates, 3 factors, 0.5y) at 2000 paths, EULER_FUNCTIONAL: L_j(t_i) windows of 16 paths, numeraire(5.0),
 *                      swaption and Bermudan prices, Bermudan regression-free exercise statistics
 */
import java.io.PrintWriter;
import java.util.HashMap;
import java.util.Map;

import net.finmath.functions.NormalDistribution;
import net.finmath.marketdata.model.curves.DiscountCurveFromForwardCurve;
import net.finmath.marketdata.model.curves.ForwardCurve;
import net.finmath.marketdata.model.curves.ForwardCurveInterpolation;
import net.finmath.montecarlo.BrownianMotion;
import net.finmath.montecarlo.BrownianMotionFromMersenneRandomNumbers;
import net.finmath.montecarlo.RandomVariableFromArrayFactory;
import net.finmath.montecarlo.assetderivativevaluation.MonteCarloBlackScholesModel;
import net.finmath.montecarlo.assetderivativevaluation.products.EuropeanOption;
import net.finmath.montecarlo.interestrate.CalibrationProduct;
import net.finmath.montecarlo.interestrate.LIBORMarketModel;
import net.finmath.montecarlo.interestrate.LIBORMonteCarloSimulationFromLIBORModel;
import net.finmath.montecarlo.interestrate.models.LIBORMarketModelFromCovarianceModel;
import net.finmath.montecarlo.interestrate.models.covariance.LIBORCorrelationModelExponentialDecay;
import net.finmath.montecarlo.interestrate.models.covariance.LIBORCovarianceModelFromVolatilityAndCorrelation;
import net.finmath.montecarlo.interestrate.models.covariance.LIBORVolatilityModelFourParameterExponentialForm;
import net.finmath.montecarlo.interestrate.products.BermudanSwaption;
import net.finmath.montecarlo.interestrate.products.Swaption;
import net.finmath.montecarlo.process.EulerSchemeFromProcessModel;
import net.finmath.randomnumbers.MersenneTwister;
import net.finmath.stochastic.RandomVariable;
import net.finmath.time.TimeDiscretization;
import net.finmath.time.TimeDiscretizationFromArray;

public class GenerateGolden {

	private static String hex(final double x) { return "\"" + Long.toHexString(Double.doubleToRawLongBits(x)) + "\""; }

	private static String hex(final double[] x, final int from, final int to) {
		final StringBuilder s = new StringBuilder("[");
		for(int i = from; i < to; i++) { s.append(i > from ? "," : "").append(hex(x[i])); }
		return s.append("]").toString();
	}

	public static void main(final String[] args) throws Exception {
		final String dir = args.length > 0 ? args[0] : ".";

		/* ---- MersenneTwister + AS241 ---- */
		try(PrintWriter w = new PrintWriter(dir + "/ref_mt.json")) {
			w.println("{\"what\": \"net.finmath.randomnumbers.MersenneTwister(seed).nextDouble(), NormalDistribution.inverseCumulativeDistribution\",");
			w.println(" \"uniforms\": {");
			final long[] seeds = { 3141, 31415, 53252, -1 };
			for(int s = 0; s < seeds.length; s++) {
				final MersenneTwister mt = new MersenneTwister(seeds[s]);
				final double[] u = new double[64];
				for(int i = 0; i < u.length; i++) { u[i] = mt.nextDouble(); }
				w.println("  \"" + seeds[s] + "\": " + hex(u, 0, u.length) + (s + 1 < seeds.length ? "," : ""));
			}
			w.println(" },");
			final MersenneTwister mt = new MersenneTwister(3141);
			final double[] z = new double[64];
			for(int i = 0; i < z.length; i++) { z[i] = NormalDistribution.inverseCumulativeDistribution(mt.nextDouble()); }
			w.println(" \"icdf_of_seed_3141\": " + hex(z, 0, z.length) + "}");
		}

		/* ---- Brownian increments ---- */
		try(PrintWriter w = new PrintWriter(dir + "/ref_brownian.json")) {
			final TimeDiscretization td = new TimeDiscretizationFromArray(0.0, 4, 0.5);
			final BrownianMotion bm = new BrownianMotionFromMersenneRandomNumbers(td, 3, 16, 3141);
			w.println("{\"what\": \"BrownianMotionFromMersenneRandomNumbers(TimeDiscretizationFromArray(0.0, 4, 0.5), 3 factors, 16 paths, seed 3141)\", \"increments\": [");
			for(int t = 0; t < 4; t++) {
				for(int f = 0; f < 3; f++) {
					final double[] x = bm.getBrownianIncrement(t, f).getRealizations();
					w.println("  " + hex(x, 0, x.length) + (t == 3 && f == 2 ? "" : ","));
				}
			}
			w.println("]}");
		}

		/* ---- C1 shape: Black-Scholes ---- */
		try(PrintWriter w = new PrintWriter(dir + "/ref_bs.json")) {
			final TimeDiscretization td = new TimeDiscretizationFromArray(0.0, 100, 0.05);
			final MonteCarloBlackScholesModel model = new MonteCarloBlackScholesModel(td, 1000, 1.0, 0.05, 0.30);      // seed 3141, EULER_FUNCTIONAL
			final double[] sT = model.getAssetValue(5.0, 0).getRealizations();
			final double price = new EuropeanOption(5.0, 1.05).getValue(model);
			w.println("{\"what\": \"MonteCarloBlackScholesModel(td(0, 100, 0.05), 1000 paths, S0 1.0, r 0.05, sigma 0.30); EuropeanOption(5.0, 1.05)\",");
			w.println(" \"asset_at_maturity_paths_0_16\": " + hex(sT, 0, 16) + ", \"call_price\": " + hex(price) + "}");
		}

		/* ---- C4 / C5 shape: LIBOR market model, 40 rates, 3 factors ---- */
		try(PrintWriter w = new PrintWriter(dir + "/ref_lmm.json")) {
			final int paths = 2000, factors = 3;
			final TimeDiscretization tenor = new TimeDiscretizationFromArray(0.0, 40, 0.5);
			final TimeDiscretization td = new TimeDiscretizationFromArray(0.0, 40, 0.5);
			final ForwardCurve forwardCurve = ForwardCurveInterpolation.createForwardCurveFromForwards("forwardCurve",
					new double[] { 0.5, 1.0, 2.0, 5.0, 40.0 }, new double[] { 0.05, 0.05, 0.05, 0.05, 0.05 }, 0.5);
			final LIBORVolatilityModelFourParameterExponentialForm vol = new LIBORVolatilityModelFourParameterExponentialForm(td, tenor, 0.2, 0.0, 0.25, 0.3, false);
			final LIBORCorrelationModelExponentialDecay corr = new LIBORCorrelationModelExponentialDecay(td, tenor, factors, 0.1);
			final LIBORCovarianceModelFromVolatilityAndCorrelation cov = new LIBORCovarianceModelFromVolatilityAndCorrelation(td, tenor, vol, corr);
			final Map<String, String> properties = new HashMap<>();
			properties.put("measure", LIBORMarketModelFromCovarianceModel.Measure.SPOT.name());
			properties.put("stateSpace", LIBORMarketModelFromCovarianceModel.StateSpace.LOGNORMAL.name());
			final LIBORMarketModel lmm = LIBORMarketModelFromCovarianceModel.of(tenor, null, forwardCurve, new DiscountCurveFromForwardCurve(forwardCurve),
					new RandomVariableFromArrayFactory(), cov, new CalibrationProduct[0], properties);
			final BrownianMotion bm = new BrownianMotionFromMersenneRandomNumbers(td, factors, paths, 3141);
			final EulerSchemeFromProcessModel process = new EulerSchemeFromProcessModel(lmm, bm, EulerSchemeFromProcessModel.Scheme.EULER_FUNCTIONAL);
			final LIBORMonteCarloSimulationFromLIBORModel sim = new LIBORMonteCarloSimulationFromLIBORModel(process);

			w.println("{\"what\": \"LMM 40 x 0.5y forwards at 5 %, vol (0.2, 0, 0.25, 0.3), correlation decay 0.1 reduced to 3 factors, SPOT, LOGNORMAL, 2000 paths, seed 3141, EULER_FUNCTIONAL\",");
			// the factor matrix the reference's PCA produced: the oracle takes it as an INPUT, so this pins the factor reduction as well
			w.println(" \"factor_loadings_t0\": [");
			for(int j = 0; j < 40; j++) {
				final RandomVariable[] fl = cov.getFactorLoading(0, j, null);
				final double[] row = new double[factors];
				for(int k = 0; k < factors; k++) { row[k] = fl[k].doubleValue(); }
				w.println("  " + hex(row, 0, factors) + (j < 39 ? "," : ""));
			}
			w.println(" ],");
			final int[][] picks = { { 1, 39 }, { 7, 8 }, { 20, 21 }, { 39, 39 }, { 40, 39 } };
			w.println(" \"libor_windows_paths_0_16\": {");
			for(int i = 0; i < picks.length; i++) {
				final double[] x = process.getProcessValue(picks[i][0], picks[i][1]).getRealizations();
				w.println("  \"" + picks[i][0] + "," + picks[i][1] + "\": " + hex(x, 0, 16) + (i + 1 < picks.length ? "," : ""));
			}
			w.println(" },");
			final double[] numeraire = sim.getNumeraire(5.0).getRealizations();
			w.println(" \"numeraire_5y_paths_0_16\": " + hex(numeraire, 0, 16) + ",");
			final double[] fixing = new double[10], payment = new double[10], swaprates = new double[10];
			for(int i = 0; i < 10; i++) { fixing[i] = 5.0 + 0.5 * i; payment[i] = 5.5 + 0.5 * i; swaprates[i] = 0.05; }
			w.println(" \"swaption_5y_into_5y_at_5pct\": " + hex(new Swaption(5.0, fixing, payment, swaprates).getValue(sim)) + ",");
			final int n = 20;
			final boolean[] isExercise = new boolean[n];
			final double[] f = new double[n], l = new double[n], p = new double[n], notional = new double[n], k = new double[n];
			for(int i = 0; i < n; i++) { isExercise[i] = true; f[i] = 5.0 + 0.5 * i; l[i] = 0.5; p[i] = 5.5 + 0.5 * i; notional[i] = 1.0; k[i] = 0.05; }
			final BermudanSwaption bermudan = new BermudanSwaption(isExercise, f, l, p, notional, k);
			w.println(" \"bermudan_20_exercise_dates\": " + hex(bermudan.getValue(sim)) + "}");
		}
		System.out.println("reference-held fixtures written to " + dir);
	}
}
