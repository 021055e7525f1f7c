"""tol_b200 -- Python test/bench harness over libtolcuda (the B200-native evaluator of tol's SNOPT
user function).  The product is the C-ABI library declared in include/tolcuda.h; this package only
loads it, moves buffers and launches one process per GPU.  Nothing here computes F or G."""
from .lib import LIB_PATH, TolcudaError, load  # noqa: F401
from .evaluator import (Evaluator, PeerBuffer, G7, S10, bounds, config_from_files, initial_guess, make_config,  # noqa: F401
                        problem_dims, problem_pattern, problem_pattern_csc, read_params, write_results_json,
                        write_results_txt)
from . import synth  # noqa: F401
