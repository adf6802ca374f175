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

THIRD-PARTY ANCHORS (round 2).  The arithmetic's real home is the go-ethereum lineage (polygon-edge's Backend: btcec-based
recovery + legacy Keccak).  tests/golden/third_party_recover.json holds published known answers from that lineage --
go-ethereum's crypto/signature_test.go triple (testmsg, testsig, testpubkey), the EIP-155 worked example, five RFC 6979
secp256k1 known answers (nonce, r, s) -- each validated mathematically (the triple recovers the PUBLISHED key; the RFC 6979
entries reproduce the published k, r, s bit for bit; OpenSSL agrees on the verify equation) and run against both oracles
(tests/test_oracle_crypto.py) and the CUDA kernels (tests/test_gpu_round2.py).  Which malformed inputs each source accepts:

  source                                   high-s   v >= 2                      x = r + n
  this engine / this oracle                accept   reject                      unsupported (reject)
  go-ethereum crypto.Ecrecover (libsecp)   accept   2,3 = x = r + n             accept when r + n < p
  btcec RecoverCompact                     accept   header 27..34 = recid 0..3  accept
  polygon-edge crypto.RecoverPubkey        accept   V != 1 is READ AS 0         (recid 2,3 unreachable)

High-s is accepted everywhere on the recover path (low-s is a transaction rule, EIP-2), and so it is here.  The two places
this engine is stricter -- V outside {0,1}, and nonce points with x >= n (probability 2^-128, not targetable) -- can only be
reached by a Byzantine sender and only make that sender's own message invalid on this node.

CPU ARMS FOR THE BENCH (all test / bench infrastructure, none on the product path).  ``c/ibft_oracle.c`` is the plain port and
THE checker (4x64-bit limbs, fixed 4-bit windows over 256 doublings, Fermat inversions, no endomorphism -- deliberately unlike
the CUDA code).  ``c/fast_recover.c`` is a tuned arm (GLV + wNAF, lazily reduced field, binary inversions; ~1.8x the plain port
per core) and ``c/ossl_recover.c`` an OpenSSL-3 arm (slower than the port on this curve); both are checked bit for bit against
the plain port in tests/test_oracle_crypto.py and against the golden bitmap in every bench run that times them.  bench.py's
``cpu_baseline.value`` and ``--impl reference`` use the fastest arm that reproduced the golden verdicts on the host.
"""
