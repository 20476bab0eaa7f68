"""TEST INFRASTRUCTURE ONLY: no-op matplotlib so the reference drivers import (they plot at module level)."""
