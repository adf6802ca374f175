package gpubackend

// COMPILE-UNVERIFIED (no Go toolchain in the authoring image).  Mirrors go-ibft_b200/host/verifier.hpp (GpuVerifier), which
// IS built and tested: same cache keys, same "malformed => false", same Prefetch batching hook.

/*
#include "ibft_verify.h"
*/
import "C"

import (
	"encoding/binary"
	"sync"

	"github.com/0xPolygon/go-ibft/messages"
	"github.com/0xPolygon/go-ibft/messages/proto"
)

// Verifier implements core.Verifier (core/backend.go:37-56) on top of the engine.  Embed it in the node's Backend; the
// remaining Backend methods (message construction, BuildProposal, InsertProposal, ...) stay as they are.
type Verifier struct {
	Eng             *Engine
	IsProposerFn    func(id []byte, height, round uint64) bool // embedder policy, not signature work
	IsValidBlockFn  func(raw []byte) bool
	mu              sync.Mutex
	cache           map[string]bool
	hashCache       map[string][32]byte
	currentHeight   uint64 // committed seals carry no height: they are checked against the running sequence's validators
}

func NewVerifier(e *Engine) *Verifier {
	return &Verifier{Eng: e, cache: map[string]bool{}, hashCache: map[string][32]byte{}}
}

func (v *Verifier) SetCurrentHeight(h uint64) { v.mu.Lock(); v.currentHeight = h; v.cache = map[string]bool{}; v.mu.Unlock() }

func (v *Verifier) IsValidProposal(raw []byte) bool { return v.IsValidBlockFn == nil || v.IsValidBlockFn(raw) }
func (v *Verifier) IsProposer(id []byte, h, r uint64) bool {
	return v.IsProposerFn != nil && v.IsProposerFn(id, h, r)
}

func senderItem(m *proto.IbftMessage) (Item, string, bool) {
	if m == nil || m.View == nil || len(m.From) != 20 || len(m.Signature) != 65 {
		return Item{}, "", false
	}
	payload, err := m.PayloadNoSig() // messages/proto/helper.go:13-27
	if err != nil {
		return Item{}, "", false
	}
	var hb [8]byte
	binary.BigEndian.PutUint64(hb[:], m.View.Height)
	key := "S" + string(hb[:]) + string(m.Signature) + string(payload)
	return Item{Kind: C.IBFT_KIND_PAYLOAD, Sig: m.Signature, Signer: m.From, Payload: payload, Height: m.View.Height}, key, true
}

func (v *Verifier) sealItem(hash []byte, seal *messages.CommittedSeal) (Item, string, bool) {
	if hash == nil || seal == nil || len(hash) != 32 || len(seal.Signer) != 20 || len(seal.Signature) != 65 {
		return Item{}, "", false
	}
	var hb [8]byte
	binary.BigEndian.PutUint64(hb[:], v.currentHeight)
	key := "C" + string(hb[:]) + string(seal.Signature) + string(seal.Signer) + string(hash)
	return Item{Kind: C.IBFT_KIND_SEAL, Sig: seal.Signature, Signer: seal.Signer, Hash: hash, Height: v.currentHeight}, key, true
}

func (v *Verifier) lookupOrVerify(it Item, key string) bool {
	v.mu.Lock()
	if ok, hit := v.cache[key]; hit {
		v.mu.Unlock()
		return ok
	}
	v.mu.Unlock()
	res, err := v.Eng.VerifyBatch([]Item{it})
	if err != nil {
		return false // launch failure: no verdict, never true
	}
	v.mu.Lock()
	v.cache[key] = res[0]
	v.mu.Unlock()
	return res[0]
}

// IsValidValidator: signer of msg.Signature over Keccak-256(PayloadNoSig) == msg.From and From is a validator at
// msg.View.Height (core/backend.go:41-45).
func (v *Verifier) IsValidValidator(m *proto.IbftMessage) bool {
	it, key, ok := senderItem(m)
	return ok && v.lookupOrVerify(it, key)
}

// IsValidCommittedSeal (core/backend.go:53-55).
func (v *Verifier) IsValidCommittedSeal(hash []byte, seal *messages.CommittedSeal) bool {
	it, key, ok := v.sealItem(hash, seal)
	return ok && v.lookupOrVerify(it, key)
}

// IsValidProposalHash with the synthetic convention of SURVEY.md §8(c): Keccak-256(Keccak-256(raw) || u64_be(round)).
// A real embedder substitutes its own block hash here; the point is that it is computed once per proposal, not per message.
func (v *Verifier) IsValidProposalHash(p *proto.Proposal, hash []byte) bool {
	if p == nil || len(hash) != 32 {
		return false
	}
	var rb [8]byte
	binary.BigEndian.PutUint64(rb[:], p.Round)
	key := string(p.RawProposal) + string(rb[:])
	v.mu.Lock()
	want, hit := v.hashCache[key]
	v.mu.Unlock()
	if !hit {
		inner, err := v.Eng.Keccak256(p.RawProposal)
		if err != nil {
			return false
		}
		want, err = v.Eng.Keccak256(append(inner[:], rb[:]...))
		if err != nil {
			return false
		}
		v.mu.Lock()
		v.hashCache[key] = want
		v.mu.Unlock()
	}
	return string(want[:]) == string(hash)
}

// Prefetch verifies, in ONE device call, every sender signature (and committed seal, and nested certificate signature) of
// the messages a handler is about to validate; the per-message methods above then answer from the cache.  Called by the
// batching shim in messages.GetValidMessages / GetExtendedRCC (see INTEGRATION.md).
func (v *Verifier) Prefetch(msgs []*proto.IbftMessage, withSeals bool) {
	var items []Item
	var keys []string
	seen := map[string]bool{}
	var visit func(m *proto.IbftMessage)
	add := func(it Item, key string, ok bool) {
		v.mu.Lock()
		_, hit := v.cache[key]
		v.mu.Unlock()
		if ok && !hit && !seen[key] {
			seen[key] = true
			items, keys = append(items, it), append(keys, key)
		}
	}
	visit = func(m *proto.IbftMessage) {
		if m == nil {
			return
		}
		add(senderItem(m))
		if withSeals {
			if seal := messages.ExtractCommittedSeal(m); seal != nil {
				add(v.sealItem(messages.ExtractCommitHash(m), seal))
			}
		}
		if pc := messages.ExtractLatestPC(m); pc != nil {
			visit(pc.ProposalMessage)
			for _, p := range pc.PrepareMessages {
				visit(p)
			}
		}
		if rcc := messages.ExtractRoundChangeCertificate(m); rcc != nil {
			for _, rc := range rcc.RoundChangeMessages {
				visit(rc)
			}
		}
	}
	for _, m := range msgs {
		visit(m)
	}
	res, err := v.Eng.VerifyBatch(items)
	if err != nil {
		return // no verdicts cached
	}
	v.mu.Lock()
	for i, k := range keys {
		v.cache[k] = res[i]
	}
	v.mu.Unlock()
}
