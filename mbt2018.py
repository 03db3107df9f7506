"""Drop-in for the `compress` / `decompress` commands of the reference's mbt2018.py (mbt2018.py:131-295): the base model's
one-shot compression with an actual bitstream (range-coded `string` / `side_string`) and its decoder.

    python mbt2018.py --num_filters 192 --checkpoint_dir ./checkpoints compress [--results_dir ./results] <runname> <input> [out.tfci]
    python mbt2018.py --num_filters 192 --checkpoint_dir ./checkpoints decompress <runname> <in.tfci> [out.png]"""
import sga_b200  # noqa: F401  (registers the package under its import name)
from sga_b200.cli import main_mbt2018

if __name__ == "__main__":
    main_mbt2018()
