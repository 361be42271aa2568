"""`python demo.py ...`: one image, same flags as the reference's demo.py (demo.py:15-76); see conzic_b200/cli.py."""
from conzic_b200.cli import demo_main

if __name__ == "__main__":
    demo_main()
