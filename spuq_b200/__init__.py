"""spuq_b200 -- B200-native stochastic-Galerkin operator application and PCG for spuq.

Drop-in surface (same names as the reference):
    Multiindex, MultiindexSet, MultiVector, MultiVectorSharedBasis, SlabMultiVector,
    MultiOperator, PreconditioningOperator, pcg, prepare_rhs, pcg_solve, inner, FlatVector, the leaf operators,
    ListCoefficientField, NormalRV, UniformRV.
"""
from .multiindex import Multiindex, MultiindexSet
from .polyquad import (LegendrePolynomials, StochasticHermitePolynomials, JacobiPolynomials,
                       ChebyshevT, ChebyshevU)
from .random_variable import NormalRV, UniformRV, RandomVariable
from .coefficient_field import CoefficientField, ListCoefficientField, ParametricCoefficientField
from .linalg import (Basis, CanonicalBasis, BasisMismatchError, Vector, FlatVector, inner, Operator,
                     MatrixOperator, DiagonalMatrixOperator, MatrixSolveOperator,
                     MultiplicationOperator, ScipyOperator, ScipySolveOperator)
from .multi_vector import MultiVector, MultiVectorSharedBasis, SlabMultiVector
from .multi_operator import MultiOperator, PreconditioningOperator, DeviceBlocks
from .pcg import pcg
from .adaptive_solver import prepare_rhs, pcg_solve, error_norm
from .sampling import compute_parametric_sample_solution, sample_solutions
from .marking import Marking
from .residual import neighbour_combination, column_energy_norms, lambda_boundary, evaluate_upper_tail_bound

__version__ = "0.1.0"
