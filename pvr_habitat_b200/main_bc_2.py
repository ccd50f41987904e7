"""Drop-in for the reference's main_bc_2.py: `run(flags)` with the reference's flags (pvr_habitat_b200.arguments).
pre-embedded observations (<data_path>/<env>_<embedding>.pickle) -> BC. The loop, file names, statistics and checkpoint
schema are in pvr_habitat_b200.bc_run; the simulator hooks (`make_environment`, `test`) are optional arguments."""
from .arguments import parser
from .bc_run import init_distributed_from_env, run_bc


def run(flags, make_environment=None, test=None, **trainer_kwargs):
    return run_bc(flags, 'bc2', make_environment=make_environment, test=test, trainer_kwargs=trainer_kwargs)


if __name__ == '__main__':
    init_distributed_from_env()  # torchrun: one rank per GPU, sequences of the global batch split over the ranks
    run(parser.parse_args())
