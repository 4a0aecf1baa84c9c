from typing import Any

from ._restated import VaeImageProcessor  # noqa: F401

PipelineImageInput = Any
