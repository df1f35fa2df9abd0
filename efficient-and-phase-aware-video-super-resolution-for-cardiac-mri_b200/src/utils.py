"""Image helpers shared by the runners (reference src/utils.py:1-20)."""

_STATS = {'acdc': (54.089, 48.084), 'dsb15': (51.193, 52.671)}


def denormalize(imgs, dataset):
    """Maps normalised frames back to 8-bit intensities: round(x * std + mean) clamped to [0, 255].
    `dataset` selects the statistics ('acdc' or 'dsb15'); the input is left untouched."""
    try:
        mean, std = _STATS[dataset]
    except KeyError:
        raise ValueError(f"The name of the dataset should be 'acdc' or 'dsb15'. Got {dataset}.") from None
    return (imgs * std + mean).round_().clamp_(0, 255)
