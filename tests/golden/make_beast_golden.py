"""Golden Beast stream from the REFERENCE PROGRAM's own network output.

Runs oracle/_ref/readsb_cpu (the stock reference program, built by `make -C oracle readsb-pair`) on a capture with
`--net --net-verbatim --net-connector 127.0.0.1,<port>,beast_out --modeac --throttle` and records every byte the program
sends to that TCP client: that is modesSendBeastOutput (net_io.c:1655-1714) at work.  --net-verbatim makes outputMessage
(net_io.c:5846-5848) forward every accepted frame (without it the first message of each aircraft is held back by the
tracker) and sends the bytes as received.  The capture starts with 1.2 s of silence so that the connection is up before
the first frame; only the traffic part is stored.  Heartbeat records (0x1a '1' + nine zero bytes) are dropped.

    python tests/golden/make_beast_golden.py
"""
import json
import socket
import subprocess
import sys
import tempfile
import threading
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
from readsb_b200 import synth  # noqa: E402

SILENCE = 2_880_000          # samples of (127, 127): no preamble, no reply
TRAFFIC = 240_000
GEN = dict(seed=77, frames_per_sec=4000.0, df_mask=synth.MODEAC | synth.DF17 | synth.DF11 | synth.AP, n_icao=6, amp=(0.35, 0.9),
           p_bit_error=0.4)


def split_records(data: bytes):
    recs, i = [], 0
    while i < len(data):
        assert data[i] == 0x1A
        j = i + 2
        need = {0x31: 2, 0x32: 7, 0x33: 14}[data[i + 1]] + 7
        got = 0
        while got < need:
            j += 2 if data[j] == 0x1A else 1
            got += 1
        recs.append(data[i:j]); i = j
    return recs


def main():
    traffic = synth.generate(TRAFFIC, **GEN)
    cap = np.concatenate([np.full(2 * SILENCE, 127, np.uint8), traffic])
    srv = socket.socket(); srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1); srv.bind(("127.0.0.1", 0)); srv.listen(1)
    port = srv.getsockname()[1]
    data = bytearray()

    def reader():
        c, _ = srv.accept(); c.settimeout(10)
        try:
            while True:
                b = c.recv(65536)
                if not b:
                    break
                data.extend(b)
        except OSError:
            pass
    t = threading.Thread(target=reader); t.start()
    with tempfile.NamedTemporaryFile(suffix=".bin") as f:
        cap.tofile(f.name)
        subprocess.run([str(ROOT / "oracle" / "_ref" / "readsb_cpu"), "--device-type", "ifile", "--ifile", f.name, "--throttle", "--quiet",
                        "--modeac", "--net", "--net-verbatim", "--net-connector", f"127.0.0.1,{port},beast_out"], check=True, timeout=120,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    t.join(timeout=15)
    recs = [r for r in split_records(bytes(data)) if r != b"\x1a1" + bytes(9)]
    stream = b"".join(recs)
    print(len(recs), "records", len(stream), "bytes; types", {chr(k): sum(1 for r in recs if r[1] == k) for k in (0x31, 0x32, 0x33)})
    np.savez_compressed(HERE / "beast_stream.npz", traffic=traffic, beast=np.frombuffer(stream, np.uint8),
                        meta=np.frombuffer(json.dumps(dict(silence_samples=SILENCE, generator=GEN, n_records=len(recs))).encode(), np.uint8))


if __name__ == "__main__":
    main()
