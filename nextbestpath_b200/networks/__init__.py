from .nbp_model import NBP  # noqa: F401
