"""Host-side scheduler arithmetic (scalars only; the tensor update runs in blend.cu).  Restates diffusers 0.28
`EulerDiscreteScheduler` with the SDXL scheduler config and `DDPMScheduler.add_noise` as used at
FaithDiff/pipelines/pipeline_FaithDiff_tlc.py:1249,1276,1336,1373."""
import numpy as np
import torch


class _Cfg(dict):
    __getattr__ = dict.__getitem__


def _alphas_cumprod(beta_start, beta_end, n):
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class EulerDiscreteScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1):
        self.config = _Cfg(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                           steps_offset=steps_offset, timestep_spacing="leading", prediction_type="epsilon")
        self.alphas_cumprod = _alphas_cumprod(beta_start, beta_end, num_train_timesteps)
        sig = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).flip(0)
        self.sigmas = torch.cat([sig, torch.zeros(1)])
        self.timesteps = torch.from_numpy(np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=float)[::-1].copy()).to(torch.float32)
        self._step_index = None

    @property
    def init_noise_sigma(self):
        return (self.sigmas.max() ** 2 + 1) ** 0.5

    def set_timesteps(self, num_inference_steps, device=None):
        n = self.config.num_train_timesteps
        ratio = n // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.float32)
        ts += self.config.steps_offset
        sig = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        sig = np.concatenate([sig, [0.0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sig).to(torch.float32)          # kept on the host, like the reference
        self.timesteps = torch.from_numpy(ts.astype(np.float32)).to(device=device)
        self.num_inference_steps = num_inference_steps
        self._step_index = None

    def step_scalars(self, i):
        """(sigma_i, sigma_{i+1}, scale_model_input divisor for step i) as the fp32 values torch computes."""
        s, sn = self.sigmas[i], self.sigmas[i + 1]
        return float(s), float(sn), float((s ** 2 + 1) ** 0.5)


class DDPMScheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
        self.config = _Cfg(num_train_timesteps=num_train_timesteps)
        self.alphas_cumprod = _alphas_cumprod(beta_start, beta_end, num_train_timesteps)

    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        t = timesteps.to(original_samples.device)
        a = (ac[t] ** 0.5).flatten()
        b = ((1 - ac[t]) ** 0.5).flatten()
        while a.ndim < original_samples.ndim:
            a, b = a.unsqueeze(-1), b.unsqueeze(-1)
        return a * original_samples + b * noise
