from .._restated import randn_tensor  # noqa: F401
