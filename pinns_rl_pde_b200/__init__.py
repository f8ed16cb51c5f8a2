"""pinns_rl_pde_b200 -- B200-native hot path of pinnrl (PINNs + RL adaptive sampling for PDEs).

MLP forward with forward-mode derivative jets, PDE residual, fused loss and the hand-written
reverse pass, behind the reference's PINNModel / PDE / PDETrainer API.  All arithmetic runs in
libpinnk.so (CUDA, sm_100a); there is no CPU or PyTorch fallback.
"""
from . import _lib
from .functional import (compute_loss, compute_residual, jets, loss_and_flat_grad, model_forward,
                         score_residual)
from .neural_networks import (Config, FeedForwardNetwork, FourierNetwork, ModelConfig, PINNModel, ResNet,
                              SIREN, make_model)
from .pdes import (AllenCahnEquation, BlackScholesEquation, BurgersEquation, CahnHilliardEquation, ConvectionEquation,
                   HeatEquation, KdVEquation, PDEBase, PDEConfig, PendulumEquation, WaveEquation, create_pde)
from .training import AdaptiveLossWeights, AdaptiveWeightsConfig, LBFGSConfig, PDETrainer, TrainingConfig
from .dropin import patch_reference
from . import rl
from .rl import DQNNetwork, dqn_forward

__all__ = ["compute_loss", "compute_residual", "jets", "loss_and_flat_grad", "model_forward", "score_residual",
           "Config", "ModelConfig", "PINNModel", "FeedForwardNetwork", "ResNet", "SIREN", "FourierNetwork",
           "make_model", "PDEConfig", "PDEBase", "HeatEquation", "BurgersEquation", "KdVEquation",
           "AllenCahnEquation", "CahnHilliardEquation", "WaveEquation", "ConvectionEquation", "BlackScholesEquation", "PendulumEquation", "create_pde", "PDETrainer", "TrainingConfig", "LBFGSConfig", "AdaptiveWeightsConfig", "AdaptiveLossWeights",
           "patch_reference", "rl", "DQNNetwork", "dqn_forward"]
