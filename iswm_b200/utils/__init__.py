from .loss import CrossEntropyLoss, FocalLoss, create_loss  # noqa: F401
