"""Federated-learning runtime: data, local training, round loop, checkpoints."""
from .runner import FederatedRunner, simulate_clients  # noqa: F401
from .trainer import LocalTrainer  # noqa: F401
