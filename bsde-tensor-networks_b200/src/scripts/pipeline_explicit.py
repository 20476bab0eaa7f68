"""`python -m src.scripts.pipeline_explicit` — kept under the reference's (misleading) file name: in the reference this
module IS the implicit scheme (fixed-point loop at src/scripts/pipeline_explicit.py:96-109)."""
from src.scripts.pipeline_implicit import main

if __name__ == "__main__":
    main()
