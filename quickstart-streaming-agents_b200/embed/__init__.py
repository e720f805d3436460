"""Embedding stage stand-ins (the reference calls Bedrock / Azure OpenAI per row; north_star stubs it)."""
from .stub import StubEmbedder, PrecomputedEmbedder  # noqa: F401
