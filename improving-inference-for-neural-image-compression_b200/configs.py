"""Module-level switches of the drop-in scripts.

The reference keeps a handful of globals in its own `configs` module that the driver scripts import lazily; this module
offers the same names so that code written against the reference finds them here.  Citations are into /root/reference.

    save_opt_record       sga.py:32,188-189,271-278  log every 10 iterations and write the opt-*.npz optimisation record
    save_reconstruction   sga.py:34,280-291          dump the reconstruction of a single-image input as recon-*.png
    write_tfci_for_eval   mbt2018.py:215-217         also write the bitstream when evaluating with mbt2018.py compress
    eval_batch_num_pixels / get_eval_batch_size      configs.py:5-9  how many images form one reference batch
    checkpoint_dir        configs.py:1               default checkpoint root of the command lines

The command-line flags --save_opt_record / --save_reconstruction (cli.py) override the two booleans for one run.
"""

# ---- optional outputs of sga.py (both off, as in the reference)
save_reconstruction = False
save_opt_record = False

# ---- evaluation: the reference cuts its input into batches holding about this many pixels -- 1e7 is ten 1000x1000
# images, 25 Kodak images (768x512) or 7 Tecnick images (1200x1200).  Both loss terms are means over such a batch, so the
# batch size enters every gradient as 1/n_ref (dist.reference_batches / dist.rank_segments).
eval_batch_num_pixels = 1e7
write_tfci_for_eval = False
checkpoint_dir = "./checkpoints"


def get_eval_batch_size(num_pixels_per_image):
    """Images per reference batch for images of the given pixel count (Python's round: half to even)."""
    return round(eval_batch_num_pixels / num_pixels_per_image)
