"""pino_locoman_b200: B200-native batched SQP inner loop behind pino-locoman's plugin surface."""
from .ocp_args import OCP_ARGS  # noqa: F401
