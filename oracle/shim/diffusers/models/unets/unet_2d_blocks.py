from ..._restated import UNetMidBlock2D, get_down_block  # noqa: F401
