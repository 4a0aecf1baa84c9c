from .._restated import DDPMScheduler, EulerDiscreteScheduler  # noqa: F401


class KarrasDiffusionSchedulers: pass
