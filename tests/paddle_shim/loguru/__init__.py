"""Stand-in for loguru (the reference's models/__init__.py imports its logger).  TEST INFRASTRUCTURE."""
import logging

logger = logging.getLogger("ppvector-ref")
