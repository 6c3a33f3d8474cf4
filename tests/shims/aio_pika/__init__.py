"""Minimal in-process stand-in for the `aio_pika` surface llmq uses.

TEST / BENCH INFRASTRUCTURE ONLY.  This image has neither `aio_pika` nor a RabbitMQ server, so the
reference's host code (ref:llmq/core/broker.py, ref:llmq/workers/base.py) is run unmodified on
top of this shim: an asyncio, single-process broker with the AMQP semantics the reference relies
on — durable named queues on the default exchange, per-channel prefetch window, one task per
delivery, round-robin between consumers, ack / reject(requeue), passive declare raising on
missing queues, purge.  A production deployment imports the real aio_pika instead.
"""
from __future__ import annotations

import asyncio
import enum
import itertools
import time
from collections import deque
from typing import Any, Callable, Deque, Dict, List, Optional

from . import abc  # noqa: F401  (aio_pika.abc import path used by the reference)

__all__ = ["connect_robust", "connect", "Message", "DeliveryMode", "reset_brokers"]


class DeliveryMode(enum.IntEnum):
    NOT_PERSISTENT = 1
    PERSISTENT = 2


class ChannelNotFoundEntity(Exception):
    pass


class Message:
    def __init__(self, body: bytes, *, delivery_mode: Any = None, message_id: Optional[str] = None,
                 headers: Optional[dict] = None, **kw):
        self.body = body
        self.delivery_mode = delivery_mode
        self.message_id = message_id
        self.headers = headers or {}
        self.timestamp = kw.get("timestamp")


class IncomingMessage(abc.AbstractIncomingMessage):
    _tags = itertools.count(1)

    def __init__(self, q: "_QueueState", msg: Message, chan: "Channel", redelivered: bool):
        self.body = msg.body
        self.message_id = msg.message_id
        self.headers = msg.headers
        self.timestamp = msg.timestamp or time.time()
        self.delivery_tag = next(self._tags)
        self.redelivered = redelivered
        self._q, self._msg, self._chan, self._done = q, msg, chan, False

    def _settle(self) -> bool:
        if self._done:
            return False
        self._done = True
        self._chan._unacked.discard(self)
        self._q.unacked -= 1
        return True

    async def ack(self, multiple: bool = False) -> None:
        if self._settle():
            self._q.broker.kick()

    async def reject(self, requeue: bool = False) -> None:
        if self._settle():
            if requeue:
                self._q.ready.appendleft((self._msg, True))
            self._q.broker.kick()

    async def nack(self, multiple: bool = False, requeue: bool = True) -> None:
        await self.reject(requeue=requeue)


class _Consumer:
    def __init__(self, chan: "Channel", cb: Callable, no_ack: bool, tag: str):
        self.chan, self.cb, self.no_ack, self.tag = chan, cb, no_ack, tag


class _QueueState:
    def __init__(self, broker: "_Broker", name: str):
        self.broker, self.name = broker, name
        self.ready: Deque = deque()
        self.consumers: List[_Consumer] = []
        self.unacked = 0
        self.rr = 0
        self.delivered = 0


class _Broker:
    def __init__(self):
        self.queues: Dict[str, _QueueState] = {}
        self._tasks: set = set()

    def queue(self, name: str, create: bool = True) -> Optional[_QueueState]:
        q = self.queues.get(name)
        if q is None and create:
            q = self.queues[name] = _QueueState(self, name)
        return q

    def kick(self) -> None:
        for q in list(self.queues.values()):
            self._dispatch(q)

    def _dispatch(self, q: _QueueState) -> None:
        while q.ready and q.consumers:
            n = len(q.consumers)
            chosen = None
            for i in range(n):
                c = q.consumers[(q.rr + i) % n]
                if c.chan.closed:
                    continue
                if c.no_ack or c.chan.prefetch == 0 or len(c.chan._unacked) < c.chan.prefetch:
                    chosen = c
                    q.rr = (q.rr + i + 1) % n
                    break
            if chosen is None:
                return
            msg, redelivered = q.ready.popleft()
            im = IncomingMessage(q, msg, chosen.chan, redelivered)
            q.delivered += 1
            if chosen.no_ack:
                im._done = True
            else:
                chosen.chan._unacked.add(im)
                q.unacked += 1
            t = asyncio.get_running_loop().create_task(self._run(chosen.cb, im))
            self._tasks.add(t)
            t.add_done_callback(self._tasks.discard)

    @staticmethod
    async def _run(cb, im):
        res = cb(im)
        if asyncio.iscoroutine(res):
            await res


_BROKERS: Dict[str, _Broker] = {}


def reset_brokers() -> None:
    _BROKERS.clear()


class _PurgeOk:
    def __init__(self, n):
        self.message_count = n


class _DeclareOk:
    def __init__(self, q: _QueueState):
        self.message_count = len(q.ready)
        self.consumer_count = len(q.consumers)


class Queue(abc.AbstractQueue):
    _ctags = itertools.count(1)

    def __init__(self, chan: "Channel", state: _QueueState):
        self.channel, self._s, self.name = chan, state, state.name
        self.declaration_result = _DeclareOk(state)

    async def consume(self, callback: Callable, no_ack: bool = False, **kw) -> str:
        tag = f"ctag-{next(self._ctags)}"
        self._s.consumers.append(_Consumer(self.channel, callback, no_ack, tag))
        self._s.broker.kick()
        return tag

    async def cancel(self, consumer_tag: str, **kw) -> None:
        self._s.consumers = [c for c in self._s.consumers if c.tag != consumer_tag]

    async def purge(self, **kw) -> _PurgeOk:
        n = len(self._s.ready)
        self._s.ready.clear()
        return _PurgeOk(n)

    async def get(self, *, no_ack: bool = False, fail: bool = True, **kw):
        if not self._s.ready:
            if fail:
                raise LookupError("queue empty")
            return None
        msg, red = self._s.ready.popleft()
        im = IncomingMessage(self._s, msg, self.channel, red)
        if no_ack:
            im._done = True
        else:
            self.channel._unacked.add(im)
            self._s.unacked += 1
        return im


class _Exchange:
    def __init__(self, chan: "Channel"):
        self._chan = chan

    async def publish(self, message: Message, routing_key: str, **kw) -> None:
        q = self._chan._broker.queue(routing_key)
        q.ready.append((message, False))
        self._chan._broker._dispatch(q)


class Channel(abc.AbstractChannel):
    def __init__(self, conn: "Connection"):
        self._conn, self._broker = conn, conn._broker
        self.prefetch = 0
        self._unacked: set = set()
        self.closed = False
        self.default_exchange = _Exchange(self)

    @property
    def is_closed(self) -> bool:
        return self.closed

    async def set_qos(self, prefetch_count: int = 0, **kw) -> None:
        self.prefetch = int(prefetch_count)
        self._broker.kick()

    async def declare_queue(self, name: Optional[str] = None, *, durable: bool = False,
                            passive: bool = False, **kw) -> Queue:
        state = self._broker.queue(name, create=not passive)
        if state is None:
            raise ChannelNotFoundEntity(f"NOT_FOUND - no queue '{name}'")
        return Queue(self, state)

    async def close(self) -> None:
        if self.closed:
            return
        self.closed = True
        for q in self._broker.queues.values():
            q.consumers = [c for c in q.consumers if c.chan is not self]
        for im in list(self._unacked):  # un-acked deliveries return to the queue (AMQP semantics)
            if im._settle():
                im._q.ready.appendleft((im._msg, True))
        self._broker.kick()


class Connection(abc.AbstractConnection):
    def __init__(self, url: str):
        self.url = url
        self._broker = _BROKERS.setdefault(url, _Broker())
        self._channels: List[Channel] = []
        self.is_closed = False

    async def channel(self, **kw) -> Channel:
        ch = Channel(self)
        self._channels.append(ch)
        return ch

    async def close(self) -> None:
        if self.is_closed:
            return
        self.is_closed = True
        for ch in self._channels:
            await ch.close()

    async def __aenter__(self):
        return self

    async def __aexit__(self, *exc):
        await self.close()


# ---- TCP mode: the same surface backed by tests/shims/aio_pika/server.py --------------------------
class _RemoteIncoming(abc.AbstractIncomingMessage):
    def __init__(self, chan: "RemoteChannel", m: dict):
        self.body, self.message_id = m["body"], m.get("mid")
        self.delivery_tag, self.redelivered = m["dtag"], m.get("red", False)
        self.timestamp, self.headers = time.time(), {}
        self._chan, self._done = chan, False

    async def ack(self, multiple: bool = False) -> None:
        if not self._done:
            self._done = True
            self._chan._send({"op": "ack", "cid": self._chan.cid, "dtag": self.delivery_tag})

    async def reject(self, requeue: bool = False) -> None:
        if not self._done:
            self._done = True
            self._chan._send({"op": "reject", "cid": self._chan.cid, "dtag": self.delivery_tag, "requeue": requeue})

    async def nack(self, multiple: bool = False, requeue: bool = True) -> None:
        await self.reject(requeue=requeue)


class RemoteQueue(abc.AbstractQueue):
    _ctags = itertools.count(1)

    def __init__(self, chan: "RemoteChannel", name: str, count: int):
        self.channel, self.name = chan, name
        self.declaration_result = _PurgeOk(count)

    async def consume(self, callback: Callable, no_ack: bool = False, **kw) -> str:
        tag = f"ctag-{id(self.channel) & 0xffff}-{next(self._ctags)}"
        self.channel._callbacks[tag] = (callback, no_ack)
        await self.channel._call({"op": "consume", "name": self.name, "ctag": tag, "no_ack": no_ack})
        return tag

    async def cancel(self, consumer_tag: str, **kw) -> None:
        await self.channel._call({"op": "cancel", "ctag": consumer_tag})
        self.channel._callbacks.pop(consumer_tag, None)

    async def purge(self, **kw) -> _PurgeOk:
        r = await self.channel._call({"op": "purge", "name": self.name})
        return _PurgeOk(r.get("count", 0))


class _RemoteExchange:
    def __init__(self, chan: "RemoteChannel"):
        self._chan = chan

    async def publish(self, message: Message, routing_key: str, **kw) -> None:
        self._chan._send({"op": "publish", "key": routing_key, "body": message.body, "mid": message.message_id})
        await self._chan._conn._drain()


class RemoteChannel(abc.AbstractChannel):
    def __init__(self, conn: "RemoteConnection", cid: int):
        self._conn, self.cid = conn, cid
        self._callbacks: Dict[str, tuple] = {}
        self.closed = False
        self.default_exchange = _RemoteExchange(self)

    @property
    def is_closed(self) -> bool:
        return self.closed

    def _send(self, m: dict) -> None:
        m.setdefault("cid", self.cid)
        self._conn._write(m)

    async def _call(self, m: dict) -> dict:
        m.setdefault("cid", self.cid)
        return await self._conn._request(m)

    async def set_qos(self, prefetch_count: int = 0, **kw) -> None:
        await self._call({"op": "qos", "n": int(prefetch_count)})

    async def declare_queue(self, name: Optional[str] = None, *, durable: bool = False,
                            passive: bool = False, **kw) -> RemoteQueue:
        r = await self._call({"op": "declare", "name": name, "passive": passive})
        if not r.get("ok"):
            raise ChannelNotFoundEntity(r.get("err", "NOT_FOUND"))
        return RemoteQueue(self, name, r.get("count", 0))

    async def close(self) -> None:
        if not self.closed:
            self.closed = True
            try:
                await self._call({"op": "close_channel"})
            except Exception:
                pass


class RemoteConnection(abc.AbstractConnection):
    def __init__(self, host: str, port: int):
        self._host, self._port = host, port
        self._reader = self._writer = None
        self._pending: Dict[int, asyncio.Future] = {}
        self._rids = itertools.count(1)
        self._cids = itertools.count(1)
        self._chans: Dict[int, RemoteChannel] = {}
        self._tasks: set = set()
        self.is_closed = False

    async def _open(self):
        from .server import read_frame  # noqa: F401

        self._reader, self._writer = await asyncio.open_connection(self._host, self._port)
        self._rx = asyncio.get_running_loop().create_task(self._read_loop())
        return self

    def _write(self, m: dict) -> None:
        from .server import write_frame

        write_frame(self._writer, m)

    async def _drain(self):
        await self._writer.drain()

    async def _request(self, m: dict) -> dict:
        rid = next(self._rids)
        m["rid"] = rid
        fut = asyncio.get_running_loop().create_future()
        self._pending[rid] = fut
        self._write(m)
        await self._writer.drain()
        return await fut

    async def _read_loop(self):
        from .server import read_frame

        try:
            while True:
                m = await read_frame(self._reader)
                if m["op"] == "reply":
                    fut = self._pending.pop(m["rid"], None)
                    if fut and not fut.done():
                        fut.set_result(m)
                elif m["op"] == "deliver":
                    ch = self._chans.get(m["cid"])
                    cb = ch._callbacks.get(m["ctag"]) if ch else None
                    if cb:
                        im = _RemoteIncoming(ch, m)
                        if cb[1]:
                            im._done = True
                        t = asyncio.get_running_loop().create_task(_Broker._run(cb[0], im))
                        self._tasks.add(t)
                        t.add_done_callback(self._tasks.discard)
        except (asyncio.IncompleteReadError, ConnectionError, asyncio.CancelledError):
            pass
        finally:
            for fut in self._pending.values():
                if not fut.done():
                    fut.set_exception(ConnectionError("shim broker connection lost"))

    async def channel(self, **kw) -> RemoteChannel:
        cid = next(self._cids)
        ch = self._chans[cid] = RemoteChannel(self, cid)
        await ch._call({"op": "open"})
        return ch

    async def close(self) -> None:
        if self.is_closed:
            return
        self.is_closed = True
        for ch in self._chans.values():
            await ch.close()
        self._rx.cancel()
        try:
            self._writer.close()
        except Exception:
            pass


async def connect_robust(url: str = "amqp://guest:guest@localhost/", **kw):
    """in-process broker by default; B200Q_SHIM_BROKER=host:port selects the TCP broker server
    (tests/shims/aio_pika/server.py) so that several processes share the queues"""
    import os

    target = os.environ.get("B200Q_SHIM_BROKER")
    if target:
        host, port = target.rsplit(":", 1)
        return await RemoteConnection(host, int(port))._open()
    return Connection(url)


connect = connect_robust
