"""Oracle = test infrastructure (CPU restatement + live third-party reference wrappers).

Never imported by the product package `whisper_context_biasing_b200`.
"""
