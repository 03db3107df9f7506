"""Drop-in for the reference's bb_sga.py (bits-back variant of SGA):

    python bb_sga.py [--verbose] --num_filters 192 --checkpoint_dir ./checkpoints compress \
        [--lambda 0.01 --sga_its 2000 --rate_its 2000 --annealing_rate 1e-3 --t0 700 --results_dir ./results] <runname> <input_file>

Multi-GPU: `python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 bb_sga.py ...`."""
import sga_b200  # noqa: F401  (registers the package under its import name)
from sga_b200.cli import main_bb

if __name__ == "__main__":
    main_bb()
