"""Import-compatible with the reference's top-level module: `from control_gen_utils import control_generate_caption`."""
from conzic_b200.control_gen_utils import (POS_sequential_generation, control_generate_caption,  # noqa: F401
                                           sentiment_sequential_generation, sentiment_shuffle_generation,
                                           set_sentiment_table)
