"""Import-compatible with the reference's top-level `utils` module (utils.py:8-74)."""
from conzic_b200.utils import create_logger, format_output, get_init_text, set_seed, update_token_mask  # noqa: F401
