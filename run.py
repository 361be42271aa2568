"""`python run.py ...` (or torchrun ... run.py ...): an image directory in batches with JSON results, same flags as
the reference's run.py (run.py:15-76); see conzic_b200/cli.py."""
from conzic_b200.cli import run_main

if __name__ == "__main__":
    run_main()
