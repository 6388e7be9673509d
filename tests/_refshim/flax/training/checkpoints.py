def restore_checkpoint(ckpt_dir, target, *a, **k):
    return target


def save_checkpoint(*a, **k):
    return None
