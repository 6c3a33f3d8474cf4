"""Socket-served version of the aio_pika stand-in's broker: one process holds the queues, any number
of worker / submitter / receiver processes connect over TCP (length-prefixed msgpack frames).
TEST / BENCH INFRASTRUCTURE ONLY — this image has no RabbitMQ; a deployment uses the real thing.

    python -m aio_pika.server --port 5673          (with tests/shims on PYTHONPATH)
    B200Q_SHIM_BROKER=127.0.0.1:5673  ...any llmq process using the stand-in...

Semantics are those the reference relies on: named queues on the default
exchange, per-channel prefetch window, round-robin between consumers, ack / reject(requeue),
un-acked deliveries return to the queue when a connection drops, passive declare of a missing
queue fails, purge.
"""
from __future__ import annotations

import argparse
import asyncio
import struct
from collections import deque
from typing import Deque, Dict, List, Optional

import msgpack


async def read_frame(reader: asyncio.StreamReader):
    hdr = await reader.readexactly(4)
    (n,) = struct.unpack("<I", hdr)
    return msgpack.unpackb(await reader.readexactly(n), raw=False)


def write_frame(writer: asyncio.StreamWriter, obj) -> None:
    data = msgpack.packb(obj, use_bin_type=True)
    writer.write(struct.pack("<I", len(data)) + data)


class _Chan:
    def __init__(self, conn: "_Conn", cid: int):
        self.conn, self.cid = conn, cid
        self.prefetch = 0
        self.unacked: Dict[int, tuple] = {}  # dtag -> (queue, body, mid)
        self.closed = False


class _Consumer:
    def __init__(self, chan: _Chan, ctag: str, no_ack: bool):
        self.chan, self.ctag, self.no_ack = chan, ctag, no_ack


class _Queue:
    def __init__(self, name):
        self.name = name
        self.ready: Deque = deque()  # (body, mid, redelivered)
        self.consumers: List[_Consumer] = []
        self.rr = 0
        self.delivered = 0


class _Conn:
    def __init__(self, server: "BrokerServer", writer):
        self.server, self.writer = server, writer
        self.chans: Dict[int, _Chan] = {}


class BrokerServer:
    def __init__(self):
        self.queues: Dict[str, _Queue] = {}
        self.dtag = 0

    def queue(self, name, create=True) -> Optional[_Queue]:
        q = self.queues.get(name)
        if q is None and create:
            q = self.queues[name] = _Queue(name)
        return q

    def dispatch(self, q: _Queue) -> None:
        while q.ready and q.consumers:
            n = len(q.consumers)
            chosen = None
            for i in range(n):
                c = q.consumers[(q.rr + i) % n]
                if c.chan.closed:
                    continue
                if c.no_ack or c.chan.prefetch == 0 or len(c.chan.unacked) < c.chan.prefetch:
                    chosen = c
                    q.rr = (q.rr + i + 1) % n
                    break
            if chosen is None:
                return
            body, mid, red = q.ready.popleft()
            self.dtag += 1
            q.delivered += 1
            if not chosen.no_ack:
                chosen.chan.unacked[self.dtag] = (q.name, body, mid)
            write_frame(chosen.chan.conn.writer, {"op": "deliver", "cid": chosen.chan.cid, "ctag": chosen.ctag,
                                                  "dtag": self.dtag, "body": body, "mid": mid, "red": red})

    def kick(self):
        for q in list(self.queues.values()):
            self.dispatch(q)

    def drop_channel(self, ch: _Chan):
        ch.closed = True
        for q in self.queues.values():
            q.consumers = [c for c in q.consumers if c.chan is not ch]
        for dtag, (qn, body, mid) in sorted(ch.unacked.items(), reverse=True):
            self.queue(qn).ready.appendleft((body, mid, True))
        ch.unacked.clear()

    async def handle(self, reader, writer):
        conn = _Conn(self, writer)
        try:
            while True:
                m = await read_frame(reader)
                op, rid = m["op"], m.get("rid")
                reply = {"op": "reply", "rid": rid, "ok": True}
                ch = conn.chans.get(m.get("cid"))
                if op == "open":
                    conn.chans[m["cid"]] = _Chan(conn, m["cid"])
                elif op == "qos":
                    ch.prefetch = int(m["n"])
                elif op == "declare":
                    q = self.queue(m["name"], create=not m.get("passive"))
                    if q is None:
                        reply.update(ok=False, err=f"NOT_FOUND - no queue '{m['name']}'")
                    else:
                        reply.update(count=len(q.ready), consumers=len(q.consumers))
                elif op == "publish":
                    self.queue(m["key"]).ready.append((m["body"], m.get("mid"), False))
                elif op == "consume":
                    self.queue(m["name"]).consumers.append(_Consumer(ch, m["ctag"], bool(m.get("no_ack"))))
                elif op == "cancel":
                    for q in self.queues.values():
                        q.consumers = [c for c in q.consumers if not (c.chan is ch and c.ctag == m["ctag"])]
                elif op == "ack":
                    ch.unacked.pop(m["dtag"], None)
                elif op == "reject":
                    item = ch.unacked.pop(m["dtag"], None)
                    if item and m.get("requeue"):
                        self.queue(item[0]).ready.appendleft((item[1], item[2], True))
                elif op == "purge":
                    q = self.queue(m["name"])
                    reply["count"] = len(q.ready)
                    q.ready.clear()
                elif op == "stats":
                    reply["queues"] = {n: {"ready": len(q.ready), "consumers": len(q.consumers), "delivered": q.delivered}
                                       for n, q in self.queues.items()}
                elif op == "close_channel":
                    if ch:
                        self.drop_channel(ch)
                if rid is not None:
                    write_frame(writer, reply)
                self.kick()
                await writer.drain()
        except (asyncio.IncompleteReadError, ConnectionError):
            pass
        finally:
            for ch in conn.chans.values():
                self.drop_channel(ch)
            self.kick()
            try:
                writer.close()
            except Exception:
                pass


async def serve(host: str, port: int, ready_file: Optional[str] = None):
    srv = BrokerServer()
    server = await asyncio.start_server(srv.handle, host, port)
    if ready_file:
        with open(ready_file, "w") as f:
            f.write(str(server.sockets[0].getsockname()[1]))
    async with server:
        await server.serve_forever()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=5673)
    ap.add_argument("--ready-file", default=None)
    a = ap.parse_args()
    asyncio.run(serve(a.host, a.port, a.ready_file))
