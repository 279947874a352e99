from . import compose, fourier_transforms, iso, loss_helpers, projections  # noqa: F401
