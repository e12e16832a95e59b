"""GPU parity tests of the MODULE-LEVEL seams (Field, Sampler, NeuRADHashEncoding as stand-alone operators
through the C ABI) against the reference's own per-stage golden values and against the fused renderer.

Written in a session without GPU access: the device code was checked on the CPU by the host emulation
(tests/test_module_seams_emul.py) and the Python glue over a fake backend (tests/test_module_glue_cpu.py); this file
sorts last so that an unexpected failure here cannot hide the verdict of the already GPU-validated suites."""
import pytest

from tests import module_seam_cases as C

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["nff_static.npz", "nff_actors.npz", "nff_sharp.npz"])
def test_field_forward_matches_reference_golden(name):
    C.field_forward_matches_reference_golden(name, "cuda")


@pytest.mark.parametrize("name", ["nff_static.npz", "nff_actors.npz"])
def test_proposal_density_and_encoding_match_reference_golden(name):
    C.proposal_density_and_encoding_match_reference_golden(name, "cuda")


@pytest.mark.parametrize("name", ["nff_static.npz", "nff_actors.npz", "nff_sharp.npz"])
def test_proposal_sampler_and_module_walk_match_reference_golden(name):
    C.proposal_sampler_and_module_walk_match_reference_golden(name, "cuda")


def test_module_operator_errors_and_empty_inputs():
    C.module_operator_errors_and_empty_inputs("cuda")


@pytest.mark.parametrize("name", ["nff_static.npz", "nff_actors.npz"])
def test_training_gradients_match_reference_golden(name):
    C.training_gradients_match_reference_golden(name, "cuda")


def test_backward_stage_operators_match_torch_autograd():
    C.backward_stage_operators_match_torch_autograd("cuda")


def test_stratified_sampling_matches_reference_golden():
    C.stratified_sampling_matches_reference_golden("cuda")


@pytest.mark.parametrize("name", ["nff_static.npz", "nff_actors.npz"])
def test_training_mode_walk_runs(name):
    C.training_mode_walk_runs(name, "cuda")
