"""Federated-learning runtime: data, local training, round loop, checkpoints."""
from .runner import FederatedRunner, classification_metrics, simulate_clients  # noqa: F401
from .trainer import LocalTrainer  # noqa: F401
