"""type-hint names the reference imports from aio_pika.abc"""


class AbstractConnection:
    pass


class AbstractChannel:
    pass


class AbstractQueue:
    pass


class AbstractIncomingMessage:
    pass


class AbstractExchange:
    pass
