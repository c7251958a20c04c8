"""TEST INFRASTRUCTURE ONLY -- empty stand-in so that the reference's
``import matplotlib`` (plot helpers, never on the hot path) succeeds."""
