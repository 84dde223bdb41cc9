from .api import (estimate_background_noise, mask_attenuated_signal, mask_impulse_noise, mask_transient_noise,
                  remove_background_noise)

# older spellings used by the reference's documentation (docs/source/data-proc-func.ipynb:111-115)
estimate_noise = estimate_background_noise
remove_noise = remove_background_noise

__all__ = ["estimate_background_noise", "remove_background_noise", "estimate_noise", "remove_noise", "mask_impulse_noise",
           "mask_transient_noise", "mask_attenuated_signal"]
