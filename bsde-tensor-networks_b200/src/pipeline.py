"""`python -m src.pipeline` — the explicit scheme entry point named in the reference's README (README.md:35)."""
from src.scripts.pipeline import main

if __name__ == "__main__":
    main()
