"""Parity suite on the HOST-EMULATED build of the CUDA sources (no GPU needed): exercises the plan
builder, the term/block tables, the kernels' index arithmetic and the Python drop-in layer.  The
same checks run on the real library in tests/test_gpu_parity.py."""
import pytest

from tests import parity_suite as ps


@pytest.fixture(autouse=True)
def _backend(emu):
    yield


def test_apply_paths_legendre():
    ps.check_apply_paths(12, 4, 2, "legendre")


def test_apply_paths_hermite_shifted():
    ps.check_apply_paths(10, 3, 3, "hermite", 0.5)


def test_apply_c1_dense_callbacks():
    ps.check_apply_c1_dense_callbacks()


def test_apply_reference_known_answer():
    ps.check_apply_reference_known_answer()


def test_apply_unstructured():
    ps.check_apply_unstructured()


def test_apply_nine_point_and_odd_shapes():
    ps.check_apply_nine_point_and_odd_shapes()


def test_apply_edge_cases():
    ps.check_apply_edge_cases()


def test_apply_m_shards():
    ps.check_apply_m_shards()


def test_apply_multiblock():
    ps.check_apply_multiblock()


def test_slab_multivector_api():
    ps.check_slab_multivector_api()


def test_blas():
    ps.check_blas()


def test_pcg_reference_test():
    ps.check_pcg_reference_test()


def test_pcg_failures():
    ps.check_pcg_failures()


@pytest.mark.parametrize("precond", ["jacobi", "identity"])
def test_pcg_sg(precond):
    ps.check_pcg_sg(precond=precond)


def test_mean_preconditioner():
    ps.check_mean_preconditioner()


def test_pcg_sg_mean():
    ps.check_pcg_sg_mean()
    ps.check_pcg_sg_mean(n=40, M=2, p=2)          # several strips: fused separator system and <rho,s>


def test_sparse_lu_preconditioner():
    ps.check_sparse_lu_preconditioner()


def test_pcg_variable_mean():
    ps.check_pcg_variable_mean()


def test_pcg_solve_and_prepare_rhs():
    ps.check_pcg_solve()


def test_row_sharded_pcg_single_rank():
    ps.check_row_sharded_pcg_single_rank()


def test_parametric_samples():
    ps.check_parametric_samples()


def test_lambda_growth():
    ps.check_lambda_growth()


def test_partitioned_mean_solver():
    ps.check_partitioned_mean_solver()


def test_neighbour_combination():
    ps.check_neighbour_combination()


def test_upper_tail_bound():
    ps.check_upper_tail_bound()


def test_refine_and_table_growth():
    ps.check_refine_and_table_growth()


def test_row_bands_manual_exchange():
    ps.check_row_bands_manual_exchange()
