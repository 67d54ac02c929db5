"""B200-native Monte-Carlo path-simulation backend behind finmath-lib's RandomVariable / BrownianMotion /
EulerSchemeFromProcessModel / MonteCarloConditionalExpectationRegression seams.

The directory name carries a hyphen (repo convention); import it through ``__graft_entry__.load_package()`` which registers
it as the module ``finmath_lib_b200``.  Layout: ``csrc/`` hand-written CUDA for sm_100a + the C ABI (include/finmath_b200.h),
``native.py`` the ctypes twin of the JNI shim, the remaining modules the host-side mirror of the reference interfaces.
"""
from . import native
from .sharding import ShardContext, LOCAL, from_environment
from .stochastic import (RandomVariable, Scalar, RandomVariableFromDoubleArray, RandomVariableCuda, RandomVariableCudaFactory)
from .montecarlo import (TimeDiscretizationFromArray, BrownianMotionCuda, BrownianMotionView, CorrelatedBrownianMotion, EulerSchemeFromProcessModel, Scheme,
                         MersenneTwister, RandomNumberGeneratorFrom1D, IndependentIncrementsFromICDF, BrownianMotionFromRandomNumberGenerator,
                         MonteCarloConditionalExpectationRegression, MonteCarloConditionalExpectationRegressionLocalizedOnDependents,
                         LinearRegression)
from .models import (BlackScholesModel, HestonModel, MonteCarloAssetModel, MonteCarloBlackScholesModel,
                     LIBORVolatilityModelFourParameterExponentialForm, LIBORCorrelationModelExponentialDecay,
                     LIBORCovarianceModelFromVolatilityAndCorrelation, LIBORMarketModelFromCovarianceModel,
                     LIBORMonteCarloSimulationFromLIBORModel, factorReduction, HullWhiteModel, ShortRateVolatilityModelAsGiven)
from .products import EuropeanOption, DigitalOption, Caplet, Swaption, BermudanSwaption, BermudanOption
from .autodiff import (RandomVariableDifferentiable, RandomVariableDifferentiableAAD, RandomVariableDifferentiableAADFactory,
                       DiracDeltaApproximationMethod)
from .calibration import (LevenbergMarquardt, OptimizerFactoryLevenbergMarquardt, RegularizationMethod, CalibrationProduct, SolverException,
                          solveLinearEquationSVD)
