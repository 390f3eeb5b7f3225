"""Writer process of the extractor: reads length-prefixed pickled tasks ``(folder, ids, embeddings, targets, pooling)`` from
stdin and writes one ``<id>_embedding_and_target.pickle`` per utterance (extract.write_item).  Runs as
``python -m loco_asr_b200._writer``; it imports numpy only (no torch, no CUDA) and exits at end of input."""
import pickle
import struct
import sys


def main():
    from loco_asr_b200.extract import write_many
    inp = sys.stdin.buffer
    n = 0
    while True:
        head = inp.read(8)
        if len(head) < 8:
            break
        (size,) = struct.unpack("<Q", head)
        folder, ids, embeddings, targets, pooling = pickle.loads(inp.read(size))
        n += write_many(folder, ids, embeddings, targets, pooling)
    return 0


if __name__ == "__main__":
    sys.exit(main())
