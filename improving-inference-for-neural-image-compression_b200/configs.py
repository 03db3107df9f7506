"""The reference's configs.py (configs.py:1-13): module-level switches read by the driver scripts."""
checkpoint_dir = "./checkpoints"

# compression / evaluation settings
write_tfci_for_eval = False          # mbt2018.py:215: write the bitstream next to the statistics
eval_batch_num_pixels = 1e7          # pixels per reference batch (10 images of 1000x1000)


def get_eval_batch_size(num_pixels_per_image):
    return round(eval_batch_num_pixels / num_pixels_per_image)


# for comparing the discretization method against others (only for the non-bits-back version): sga.py:188-189, 271-278
save_opt_record = False
# sga.py:34, 280-291: dump the reconstruction of a single-image input as PNG
save_reconstruction = False
