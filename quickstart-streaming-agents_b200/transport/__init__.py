"""Topic transports.  ``filelog`` is the self-contained one (Kafka semantics on a directory)."""
from .filelog import Broker, Consumer, Message, Producer, TopicPartition  # noqa: F401
