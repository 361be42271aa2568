"""Import-compatible with the reference's top-level module of the same name: `from gen_utils import generate_caption`."""
from conzic_b200.gen_utils import *  # noqa: F401,F403
from conzic_b200.gen_utils import generate_caption, generate_caption_step  # noqa: F401
