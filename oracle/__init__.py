"""CPU oracle for the go-ibft message-verification hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the shipped product path: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import or execute it, and only as the checker / the timed CPU baseline.

PARITY STATUS: **parity unpinned by the reference.**  go-ibft (reference @ e81a63ff) contains no
cryptography: ``core.Verifier`` (core/backend.go:37-56) is an interface whose production
implementation lives in the embedding node, and the reference's tests mock it with always-true /
byte-equality closures (core/mock_test.go:105-151).  There is therefore no golden signature, hash
or address in the reference tree to pin this oracle against.  The crypto conventions restated here
are the ones SURVEY.md §8(c) fixes ([EXTERNAL]: secp256k1 per SEC 2 v2 §2.4.1, original
Keccak-256, 65-byte R||S||V signatures with V in {0,1}, address = Keccak-256(X||Y)[12:]); they are
anchored on independent known-answer vectors (Keccak KATs, privkey 1 -> G -> address
0x7e5f4552091a69125d5dfcb7b8c2659029395bdf) and cross-checked against the ``cryptography``
package (OpenSSL) in tests/test_oracle_crypto.py.  What the reference's tests DO pin -- quorum
arithmetic, validPC / validateProposal decision tables, store pruning semantics, proto encoding --
is restated in ``ibft_logic.py`` / ``ibft_proto.py`` and checked against those tables.
"""
