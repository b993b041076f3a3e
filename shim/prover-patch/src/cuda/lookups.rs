//! `impl stwo::core::channel::Channel` over the library's Blake2sChannel so that the reference's
//! `C::draw_lookup_elements(&mut lookup_elements, channel, config)` (machine.rs:239-240, `relation!`'s `LookupElements::draw`)
//! runs unchanged against the transcript the GPU prover continues.
use nexus_b200::Channel as NbChannel;
use stwo::core::channel::Channel;
use stwo::core::fields::qm31::SecureField;

pub struct ChannelAdapter<'a>(pub &'a mut NbChannel);

impl Channel for ChannelAdapter<'_> {
    const BYTES_PER_HASH: usize = 32;
    fn trailing_zeros(&self) -> u32 { let d = self.0.digest(); u128::from_le_bytes(d[..16].try_into().unwrap()).trailing_zeros() }
    fn mix_felts(&mut self, felts: &[SecureField]) { self.0.mix_felts(felts) }
    fn mix_u32s(&mut self, _data: &[u32]) { unimplemented!("not used between the tree-1 commit and the interaction trace") }
    fn mix_u64(&mut self, value: u64) { self.0.mix_u64(value) }
    fn draw_felt(&mut self) -> SecureField { self.0.draw_felts(1)[0] }
    fn draw_felts(&mut self, n_felts: usize) -> Vec<SecureField> { self.0.draw_felts(n_felts) }
    fn draw_random_bytes(&mut self) -> Vec<u8> { unimplemented!("queries are drawn inside nb200_prove") }
}
