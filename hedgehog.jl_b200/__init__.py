"""hedgehog.jl_b200 — B200-native Monte Carlo pricing path with Hedgehog.jl's solve(problem, method) API.

The directory name contains a dot, so import it through the repo-root shim: `import hedgehog_jl_b200 as hh`.
All numerics run in libhedgehog_mc.so (hand-written CUDA for sm_100a, C ABI in include/hedgehog_mc.h).
"""
from ._abi import HedgehogB200Error, load_library, LIB_PATH  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import (B200MonteCarlo, BasketPricingProblem, corr_factor, df, solve, to_ticks, yearfrac, add_yearfrac,  # noqa: F401
                  zero_rate, get_vol)
from .engine import CudaEngine, SimSpec, default_engine  # noqa: F401
from .greeks import (BatchGreekProblem, FDBackward, FDCentral, FDForward, FieldLens, FiniteDifference, ForwardAD,  # noqa: F401
                     GreekProblem, GreekResult, SecondOrderGreekProblem, SpotLens, VolLens, ZeroRateSpineLens, optic,
                     strike_grid_greeks)
from .greeks import set as set_lens  # noqa: F401
from .calibration import CalibrationProblem, CalibrationResult, OptimizerAlgo, RootFinderAlgo, basket_prices_and_jacobian  # noqa: F401,E402
from .pathdep import (ArithmeticAverage, AsianOption, BlackScholesControlVariate, AssetOrNothing, BarrierOption, CashOrNothing, DigitalOption, Down,  # noqa: F401,E402
                      GeometricAverage, GeometricControlVariate, KnockIn, KnockOut, Monitoring, Up)
